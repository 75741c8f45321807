// oracle/restir_rearch.inl — TEST INFRASTRUCTURE (CPU oracle), included by render.cpp.
//
// CPU restatement of the "rearchitected" ReSTIR DI renderer (Wyman & Panteleev 2021 style):
//   performLightPreSampling        restir_di/gpu_kernels/per_pixel_ris.cu:6-40
//   performPerPixelRIS             restir_di/gpu_kernels/per_pixel_ris.cu:44-128
//   traceShadowRays<T,S,U>         restir_di/gpu_kernels/optix_restir_di_rearch_kernels.cu:14-225
//   computeMISWeight<type,T,S>     :263-400  (useMIS_RIS = true, :10)
//   shadeAndResample<T,S>          :402-664
//   host sequence                  restir_di/restir_di_main.cpp:2423-2493; presampling RNGs :1210-1222
//                                  (128 subsets x 1024 lights, mt19937_64(894213312210))
// SampleVisibility bits (restir_di_shared.h:146-164): 0 newSample, 1 newSampleOnTemporal, 2 newSampleOnSpatiotemporal,
// 3 temporalPassedHeuristic, 4 temporalSample, 5 temporalSampleOnCurrent, 6 temporalSampleOnSpatiotemporal,
// 7 spatiotemporalPassedHeuristic, 8 spatiotemporalSample, 9 spatiotemporalSampleOnCurrent,
// 10 spatiotemporalSampleOnTemporal, 11 selectedSample.
// performPerPixelRIS runs in 8x8 blocks (cudau::dim3(8, 8) at module load): the thread at the tile origin draws the
// tile's light subset from ITS pixel RNG before anything else (and keeps the advanced state only if it is a hit).

static constexpr uint32_t kNumLightSubsets = 128;   // restir_di_shared.h:8
static constexpr uint32_t kLightSubsetSize = 1024;  // restir_di_shared.h:9
static constexpr uint32_t kNumPreSampledLights = kNumLightSubsets * kLightSubsetSize;

enum : uint32_t {
    SV_NEW = 1u << 0, SV_NEW_ON_T = 1u << 1, SV_NEW_ON_ST = 1u << 2,
    SV_T_PASSED = 1u << 3, SV_T = 1u << 4, SV_T_ON_CUR = 1u << 5, SV_T_ON_ST = 1u << 6,
    SV_ST_PASSED = 1u << 7, SV_ST = 1u << 8, SV_ST_ON_CUR = 1u << 9, SV_ST_ON_T = 1u << 10,
    SV_SELECTED = 1u << 11
};

struct PreSampledLight { // restir_di_shared.h PreSampledLight (44 B), padded to 48 B
    float emittance[3], areaPDensity;
    float position[3]; uint32_t atInfinity;
    float normal[3], pad;
};
static_assert(sizeof(PreSampledLight) == 48, "PreSampledLight record is 48 bytes");

struct orc_rearch {
    std::vector<PreSampledLight> preSampledLights;
    std::vector<uint64_t> rngs;
    std::vector<uint32_t> sampleVis[2];
};
static void rearchDestroy(orc_rearch* r) { delete r; }

static orc_rearch* rearchState(orc_frame* f) {
    if (f->rearch)
        return f->rearch;
    orc_rearch* r = new orc_rearch();
    PreSampledLight empty;
    std::memset(&empty, 0, sizeof(empty));
    r->preSampledLights.assign(kNumPreSampledLights, empty);
    r->rngs.resize(kNumPreSampledLights);
    std::mt19937_64 rngSeed(894213312210ull);
    for (auto &s : r->rngs)
        s = rngSeed();
    for (int i = 0; i < 2; ++i)
        r->sampleVis[i].assign((size_t)f->W * f->H, 0u);
    f->rearch = r;
    return r;
}

static void* rearchBufferPtr(orc_frame* f, int id, uint32_t index, size_t* bytes) {
    orc_rearch* r = rearchState(f);
    void* ptr = nullptr;
    size_t b = 0;
    switch (id) {
    case GFX_BUF_SAMPLE_VISIBILITY: ptr = r->sampleVis[index & 1].data(); b = r->sampleVis[index & 1].size() * 4; break;
    case GFX_BUF_PRESAMPLED_LIGHTS: ptr = r->preSampledLights.data(); b = r->preSampledLights.size() * 48; break;
    case GFX_BUF_PRESAMPLE_RNG: ptr = r->rngs.data(); b = r->rngs.size() * 8; break;
    default: break;
    }
    if (bytes) *bytes = b;
    return ptr;
}

static inline LightSample lightSampleOf(const PreSampledLight &l) {
    LightSample s;
    s.emittance = float3(l.emittance[0], l.emittance[1], l.emittance[2]);
    s.position = float3(l.position[0], l.position[1], l.position[2]);
    s.normal = float3(l.normal[0], l.normal[1], l.normal[2]);
    s.atInfinity = l.atInfinity;
    return s;
}

// per_pixel_ris.cu:6-40
static void performLightPreSampling(orc_frame* f, const GfxFrameParams* p, int numThreads) {
    const orc_scene* s = f->scene;
    orc_rearch* r = rearchState(f);
#pragma omp parallel for schedule(static) num_threads(numThreads)
    for (int64_t i = 0; i < (int64_t)kNumPreSampledLights; ++i) {
        PCG32RNG rng{ r->rngs[i] };
        LightSample ls;
        ls.emittance = float3(0.0f);
        ls.position = float3(0.0f);
        ls.normal = float3(0.0f);
        ls.atInfinity = 0;
        float areaPDensity = 0.0f;
        // :12-28: the first probToSampleEnvLight * lightSubsetSize lights of every subset come from the environment
        float probToSampleCurLightType = 1.0f;
        bool sampleEnvLight = false;
        if (useEnvLight(s, p)) {
            if (s->instIntegral > 0.0f) {
                const uint32_t indexInSubset = (uint32_t)i % kLightSubsetSize;
                sampleEnvLight = indexInSubset < kProbToSampleEnvLight * kLightSubsetSize;
                probToSampleCurLightType = sampleEnvLight ? kProbToSampleEnvLight : (1 - kProbToSampleEnvLight);
            }
            else {
                sampleEnvLight = true;
            }
        }
        const float ul = rng.getFloat0cTo1o();
        const float u0 = rng.getFloat0cTo1o();
        const float u1 = rng.getFloat0cTo1o();
        sampleLight(s, p, ul, sampleEnvLight, u0, u1, &ls, &areaPDensity);
        PreSampledLight &o = r->preSampledLights[i];
        o.emittance[0] = ls.emittance.x; o.emittance[1] = ls.emittance.y; o.emittance[2] = ls.emittance.z;
        o.areaPDensity = areaPDensity * probToSampleCurLightType;
        o.position[0] = ls.position.x; o.position[1] = ls.position.y; o.position[2] = ls.position.z;
        o.atInfinity = ls.atInfinity;
        o.normal[0] = ls.normal.x; o.normal[1] = ls.normal.y; o.normal[2] = ls.normal.z;
        o.pad = 0.0f;
        r->rngs[i] = rng.state;
    }
}

struct RearchShadingPoint {
    float3 positionInWorld, vOutLocal;
    ReferenceFrame shadingFrame;
    BSDF bsdf;
};
// the prologue shared by performPerPixelRIS / shadeAndResample / computeMISWeight's neighbour reconstruction
static inline RearchShadingPoint reconstructShadingPoint(const orc_frame* f, const orc_scene* s, uint32_t bufIdx, size_t pix,
                                                         const float3 &cameraPosition) {
    const GB2 gb2 = f->gb2[bufIdx][pix];
    const GB3 gb3 = f->gb3[bufIdx][pix];
    RearchShadingPoint sp;
    float3 positionInWorld(gb2.px, gb2.py, gb2.pz);
    const float3 geometricNormalInWorld = decodeVector(gb2.qGeometricNormal);
    const float3 vOut = normalize(cameraPosition - positionInWorld);
    const float frontHit = dot(vOut, geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
    sp.positionInWorld = offsetRayOrigin(positionInWorld, frontHit * geometricNormalInWorld);
    sp.shadingFrame = ReferenceFrame(decodeVector(gb3.qShadingNormal), decodeVector(gb3.qShadingTangent));
    sp.vOutLocal = sp.shadingFrame.toLocal(vOut);
    sp.bsdf = setupBsdf(s, gb3.matSlot, decodeTexCoords(gb3.qTexCoord));
    return sp;
}

// per_pixel_ris.cu:44-128
static void performPerPixelRIS(orc_frame* f, const GfxFrameParams* p, const Camera &camera, const std::vector<uint64_t> &rngIn,
                               uint32_t x, uint32_t y) {
    const orc_scene* s = f->scene;
    orc_rearch* r = rearchState(f);
    const size_t pix = (size_t)y * f->W + x;
    const uint32_t curBufIdx = p->bufferIndex & 1;

    // the tile's light subset: drawn by the thread at the tile origin from its own RNG
    const uint32_t tx = x & ~7u, ty = y & ~7u;
    uint32_t perTileLightSubsetIndex;
    {
        PCG32RNG tileRng{ rngIn[(size_t)ty * f->W + tx] };
        perTileLightSubsetIndex = std::min(dm_f2uint(tileRng.getFloat0cTo1o() * kNumLightSubsets), kNumLightSubsets - 1);
    }
    PCG32RNG rng{ rngIn[pix] };
    if (x == tx && y == ty)
        (void)rng.getFloat0cTo1o();
    const PreSampledLight* lightSubSet = &r->preSampledLights[(size_t)perTileLightSubsetIndex * kLightSubsetSize];

    if (f->gb0[curBufIdx][pix].instSlot == 0xFFFFFFFFu)
        return;
    const RearchShadingPoint sp = reconstructShadingPoint(f, s, curBufIdx, pix, camera.position);

    const uint32_t curResIndex = p->currentReservoirIndex & 1;
    Reservoir reservoir;
    reservoir.initialize(emptyLightSample());
    float selectedTargetDensity = 0.0f;
    const uint32_t numCandidates = 1u << p->log2NumCandidateSamples;
    for (uint32_t i = 0; i < numCandidates; ++i) {
        const uint32_t lightIndex = std::min(dm_f2uint(rng.getFloat0cTo1o() * kLightSubsetSize), kLightSubsetSize - 1);
        const PreSampledLight &preSampledLight = lightSubSet[lightIndex];
        const LightSample ls = lightSampleOf(preSampledLight);
        const float3 cont = performDirectLighting<false>(s, sp.positionInWorld, sp.vOutLocal, sp.shadingFrame, sp.bsdf, ls);
        const float targetDensity = convertToWeight(cont);
        const float weight = targetDensity / preSampledLight.areaPDensity;
        if (reservoir.update(ls, weight, rng.getFloat0cTo1o()))
            selectedTargetDensity = targetDensity;
    }
    float recPDFEstimate = reservoir.sumWeights / (selectedTargetDensity * reservoir.streamLength);
    if (!std::isfinite(recPDFEstimate)) {
        recPDFEstimate = 0.0f;
        selectedTargetDensity = 0.0f;
    }
    f->rng[pix] = rng.state;
    storeReservoir(f, curResIndex, pix, reservoir);
    f->reservoirInfo[curResIndex][pix] = GB1{ recPDFEstimate, selectedTargetDensity };
}

// the (tNbCoord, stNbCoord) computation shared by traceShadowRays and shadeAndResample
static inline void temporalNeighborCoord(const orc_frame* f, uint32_t curBufIdx, size_t pix, uint32_t x, uint32_t y, int* nbx, int* nby) {
    const GB1 gb1 = f->gb1[curBufIdx][pix];
    *nbx = dm_f2int(x + 0.5f - gb1.mvx);
    *nby = dm_f2int(y + 0.5f - gb1.mvy);
}
static inline void spatialNeighborCoord(const orc_frame* f, const GfxFrameParams* p, uint32_t x, uint32_t y, PCG32RNG &rng,
                                        int* nbx, int* nby, float* deltaX, float* deltaY) {
    float radius = p->spatialNeighborRadius;
    if (p->useLowDiscrepancyNeighbors) {
        const uint32_t deltaIndex = p->spatialNeighborBaseIndex + 5 * x + 7 * y;
        const float2 delta = f->neighborDeltas[deltaIndex % 1024];
        *deltaX = radius * delta.x;
        *deltaY = radius * delta.y;
    }
    else {
        radius *= std::sqrt(rng.getFloat0cTo1o());
        const float angle = 2 * kPi * rng.getFloat0cTo1o();
        float sa, ca;
        dm_sincos(angle, &sa, &ca);
        *deltaX = radius * ca;
        *deltaY = radius * sa;
    }
    *nbx = dm_f2int(x + 0.5f + *deltaX);
    *nby = dm_f2int(y + 0.5f + *deltaY);
}

// optix_restir_di_rearch_kernels.cu:14-225
template <bool withTemporalRIS, bool withSpatialRIS, bool useUnbiasedEstimator>
static void traceShadowRays(orc_frame* f, const GfxFrameParams* p, const Camera &camera, const Camera &prevCamera, uint32_t x, uint32_t y,
                            uint64_t* rayCount) {
    const orc_scene* s = f->scene;
    orc_rearch* r = rearchState(f);
    const size_t pix = (size_t)y * f->W + x;
    const uint32_t curBufIdx = p->bufferIndex & 1, prevBufIdx = (curBufIdx + 1) % 2;
    const uint32_t curResIndex = p->currentReservoirIndex & 1, prevResIndex = (curResIndex + 1) % 2;
    if (f->gb0[curBufIdx][pix].instSlot == 0xFFFFFFFFu)
        return;
    const GB2 gb2 = f->gb2[curBufIdx][pix];
    const GB3 gb3 = f->gb3[curBufIdx][pix];
    float3 positionInWorld(gb2.px, gb2.py, gb2.pz);
    const float3 geometricNormalInWorld = decodeVector(gb2.qGeometricNormal);
    const float3 shadingNormalInWorld = decodeVector(gb3.qShadingNormal);
    const float3 vOut = camera.position - positionInWorld;
    const float frontHit = dot(vOut, geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
    positionInWorld = offsetRayOrigin(positionInWorld, frontHit * geometricNormalInWorld);
    const float dist = length(vOut);

    auto visible = [&](const float3 &origin, const LightSample &ls) -> uint32_t {
        ++*rayCount;
        return evaluateVisibility(s, origin, ls) ? 1u : 0u;
    };
    auto neighborOrigin = [&](size_t nbPix) {
        const GB2 nbGb2 = f->gb2[prevBufIdx][nbPix];
        const float3 nbPositionInWorld(nbGb2.px, nbGb2.py, nbGb2.pz);
        const float3 nbGeometricNormalInWorld = decodeVector(nbGb2.qGeometricNormal);
        const float3 nbVOut = prevCamera.position - nbPositionInWorld;
        const float nbFrontHit = dot(nbVOut, nbGeometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
        return offsetRayOrigin(nbPositionInWorld, nbFrontHit * nbGeometricNormalInWorld);
    };

    uint32_t sampleVis = 0;
    LightSample newSample;
    bool newSampleIsValid;
    {
        const Reservoir reservoir = loadReservoir(f, curResIndex, pix);
        newSample = reservoir.sample;
        newSampleIsValid = reservoir.sumWeights > 0.0f;
        if (newSampleIsValid && visible(positionInWorld, newSample))
            sampleVis |= SV_NEW;
    }

    int tNbX = 0, tNbY = 0;
    float3 tNbPositionInWorld(0.0f);
    bool temporalSampleIsValid = false;
    LightSample temporalSample = emptyLightSample();
    if (withTemporalRIS) {
        temporalNeighborCoord(f, curBufIdx, pix, x, y, &tNbX, &tNbY);
        if (testNeighbor<true>(f, camera, prevBufIdx, tNbX, tNbY, dist, shadingNormalInWorld))
            sampleVis |= SV_T_PASSED;
        if (sampleVis & SV_T_PASSED) {
            const size_t nbPix = (size_t)tNbY * f->W + tNbX;
            if (p->reuseVisibilityForTemporal && !useUnbiasedEstimator) {
                if (r->sampleVis[prevBufIdx][nbPix] & SV_SELECTED)
                    sampleVis |= SV_T;
            }
            else {
                const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
                temporalSample = neighbor.sample;
                temporalSampleIsValid = neighbor.sumWeights > 0.0f;
                if (temporalSampleIsValid && visible(positionInWorld, temporalSample))
                    sampleVis |= SV_T;
            }
            if (useUnbiasedEstimator) {
                tNbPositionInWorld = neighborOrigin(nbPix);
                if (newSampleIsValid && visible(tNbPositionInWorld, newSample))
                    sampleVis |= SV_NEW_ON_T;
                if (temporalSampleIsValid && visible(positionInWorld, temporalSample))
                    sampleVis |= SV_T_ON_CUR;
            }
        }
    }

    int stNbX = 0, stNbY = 0;
    float3 stNbPositionInWorld(0.0f);
    bool spatiotemporalSampleIsValid = false;
    LightSample spatiotemporalSample = emptyLightSample();
    if (withSpatialRIS) {
        float deltaX, deltaY;
        PCG32RNG rng{ f->rng[pix] }; // the advanced state is NOT stored (:148-150)
        spatialNeighborCoord(f, p, x, y, rng, &stNbX, &stNbY, &deltaX, &deltaY);
        bool passed = testNeighbor<true>(f, camera, prevBufIdx, stNbX, stNbY, dist, shadingNormalInWorld);
        passed = passed && (stNbX != (int)x || stNbY != (int)y);
        if (passed)
            sampleVis |= SV_ST_PASSED;
        if (passed) {
            const size_t nbPix = (size_t)stNbY * f->W + stNbX;
            bool reused = false;
            if (p->reuseVisibilityForSpatiotemporal && !useUnbiasedEstimator) {
                const float threshold2 = pow2(p->radiusThresholdForSpatialVisReuse);
                const float dist2 = pow2(deltaX) + pow2(deltaY);
                reused = dist2 < threshold2;
            }
            if (reused) {
                if (r->sampleVis[prevBufIdx][nbPix] & SV_SELECTED)
                    sampleVis |= SV_ST;
            }
            else {
                const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
                spatiotemporalSample = neighbor.sample;
                spatiotemporalSampleIsValid = neighbor.sumWeights > 0.0f;
                if (spatiotemporalSampleIsValid && visible(positionInWorld, spatiotemporalSample))
                    sampleVis |= SV_ST;
            }
            if (useUnbiasedEstimator) {
                stNbPositionInWorld = neighborOrigin(nbPix);
                if (newSampleIsValid && visible(stNbPositionInWorld, newSample))
                    sampleVis |= SV_NEW_ON_ST;
                if (spatiotemporalSampleIsValid && visible(positionInWorld, spatiotemporalSample))
                    sampleVis |= SV_ST_ON_CUR;
            }
        }
    }

    if (useUnbiasedEstimator && withTemporalRIS && withSpatialRIS) {
        if ((sampleVis & SV_T_PASSED) && (sampleVis & SV_ST_PASSED)) {
            if (temporalSampleIsValid) {
                const Reservoir tNeighbor = loadReservoir(f, prevResIndex, (size_t)tNbY * f->W + tNbX);
                if (visible(stNbPositionInWorld, tNeighbor.sample))
                    sampleVis |= SV_T_ON_ST;
            }
            if (spatiotemporalSampleIsValid) {
                const Reservoir stNeighbor = loadReservoir(f, prevResIndex, (size_t)stNbY * f->W + stNbX);
                if (visible(tNbPositionInWorld, stNeighbor.sample))
                    sampleVis |= SV_ST_ON_T;
            }
        }
    }
    r->sampleVis[curBufIdx][pix] = sampleVis;
}

enum class SampleType { New = 0, Temporal, Spatiotemporal };

// optix_restir_di_rearch_kernels.cu:263-400 with useMIS_RIS = true
template <SampleType sampleType, bool withTemporalRIS, bool withSpatialRIS>
static float computeMISWeight(const orc_frame* f, const GfxFrameParams* p, const Camera &prevCamera, uint32_t prevBufIdx,
                              uint32_t prevResIndex, uint32_t maxPrevStreamLength, uint32_t sampleVis, uint32_t selfStreamLength,
                              const RearchShadingPoint &sp, int tNbX, int tNbY, int stNbX, int stNbY, uint32_t streamLength,
                              const LightSample &lightSample, float sampleTargetDensity) {
    const orc_scene* s = f->scene;
    const float numMisWeight = sampleTargetDensity;
    float denomMisWeight = numMisWeight * streamLength;

    if (sampleType != SampleType::New) {
        const float3 cont = performDirectLighting<false>(s, sp.positionInWorld, sp.vOutLocal, sp.shadingFrame, sp.bsdf, lightSample);
        float targetDensity = convertToWeight(cont);
        if (p->useUnbiasedEstimator)
            targetDensity *= sampleType == SampleType::Temporal ? ((sampleVis & SV_T_ON_CUR) ? 1 : 0) : ((sampleVis & SV_ST_ON_CUR) ? 1 : 0);
        denomMisWeight += targetDensity * selfStreamLength;
    }
    if (sampleType != SampleType::Temporal && withTemporalRIS) {
        if (sampleVis & SV_T_PASSED) {
            const size_t nbPix = (size_t)tNbY * f->W + tNbX;
            const RearchShadingPoint nb = reconstructShadingPoint(f, s, prevBufIdx, nbPix, prevCamera.position);
            const float3 cont = performDirectLighting<false>(s, nb.positionInWorld, nb.vOutLocal, nb.shadingFrame, nb.bsdf, lightSample);
            float nbTargetDensity = convertToWeight(cont);
            if (p->useUnbiasedEstimator)
                nbTargetDensity *= sampleType == SampleType::New ? ((sampleVis & SV_NEW_ON_T) ? 1 : 0) : ((sampleVis & SV_ST_ON_T) ? 1 : 0);
            const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
            const uint32_t nbStreamLength = std::min(neighbor.streamLength, maxPrevStreamLength);
            denomMisWeight += nbTargetDensity * nbStreamLength;
        }
    }
    if (sampleType != SampleType::Spatiotemporal && withSpatialRIS) {
        if (sampleVis & SV_ST_PASSED) {
            const size_t nbPix = (size_t)stNbY * f->W + stNbX;
            const RearchShadingPoint nb = reconstructShadingPoint(f, s, prevBufIdx, nbPix, prevCamera.position);
            const float3 cont = performDirectLighting<false>(s, nb.positionInWorld, nb.vOutLocal, nb.shadingFrame, nb.bsdf, lightSample);
            float nbTargetDensity = convertToWeight(cont);
            if (p->useUnbiasedEstimator)
                nbTargetDensity *= sampleType == SampleType::New ? ((sampleVis & SV_NEW_ON_ST) ? 1 : 0) : ((sampleVis & SV_T_ON_ST) ? 1 : 0);
            const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
            const uint32_t nbStreamLength = std::min(neighbor.streamLength, maxPrevStreamLength);
            denomMisWeight += nbTargetDensity * nbStreamLength;
        }
    }
    return numMisWeight / denomMisWeight;
}

// optix_restir_di_rearch_kernels.cu:402-664
template <bool withTemporalRIS, bool withSpatialRIS>
static void shadeAndResample(orc_frame* f, const GfxFrameParams* p, const Camera &camera, const Camera &prevCamera, uint32_t x, uint32_t y) {
    const orc_scene* s = f->scene;
    orc_rearch* r = rearchState(f);
    const size_t pix = (size_t)y * f->W + x;
    const uint32_t curBufIdx = p->bufferIndex & 1, prevBufIdx = (curBufIdx + 1) % 2;
    const uint32_t curResIndex = p->currentReservoirIndex & 1, prevResIndex = (curResIndex + 1) % 2;
    const GB0 gb0 = f->gb0[curBufIdx][pix];
    const GB3 gb3 = f->gb3[curBufIdx][pix];

    float3 contribution(0.01f, 0.01f, 0.01f);
    if (gb0.instSlot != 0xFFFFFFFFu) {
        PCG32RNG rng{ f->rng[pix] };
        int tNbX = 0, tNbY = 0, stNbX = 0, stNbY = 0;
        if (withTemporalRIS)
            temporalNeighborCoord(f, curBufIdx, pix, x, y, &tNbX, &tNbY);
        if (withSpatialRIS) {
            float deltaX, deltaY;
            spatialNeighborCoord(f, p, x, y, rng, &stNbX, &stNbY, &deltaX, &deltaY);
        }
        const RearchShadingPoint sp = reconstructShadingPoint(f, s, curBufIdx, pix, camera.position);
        const GfxMaterialDesc &mat = s->materials[gb3.matSlot];

        contribution = float3(0.0f);
        if (sp.vOutLocal.z > 0) {
            float3 emittance(0.0f);
            if (mat.hasEmittance)
                emittance = float3(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
            contribution += emittance / kPi;
        }

        uint32_t sampleVis = r->sampleVis[curBufIdx][pix];
        float selectedTargetDensity = 0.0f;
        Reservoir combinedReservoir;
        uint32_t combinedStreamLength = 0;
        combinedReservoir.initialize(emptyLightSample());
        float3 directCont(0.0f);
        float selectedMisWeight = 0.0f;

        const Reservoir selfRes = loadReservoir(f, curResIndex, pix);
        const GB1 selfResInfo = f->reservoirInfo[curResIndex][pix];
        const uint32_t selfStreamLength = selfRes.streamLength;
        const uint32_t maxPrevStreamLength = 20 * selfStreamLength;
        auto setSelected = [&](bool v) { sampleVis = v ? (sampleVis | SV_SELECTED) : (sampleVis & ~SV_SELECTED); };

        // new sample of the current pixel
        {
            if (selfResInfo.mvx > 0.0f && (sampleVis & SV_NEW)) {
                const LightSample lightSample = selfRes.sample;
                const float3 cont = performDirectLighting<false>(s, sp.positionInWorld, sp.vOutLocal, sp.shadingFrame, sp.bsdf, lightSample);
                const float targetDensity = convertToWeight(cont);
                float misWeight;
                if (withTemporalRIS || withSpatialRIS)
                    misWeight = computeMISWeight<SampleType::New, withTemporalRIS, withSpatialRIS>(
                        f, p, prevCamera, prevBufIdx, prevResIndex, maxPrevStreamLength, sampleVis, selfStreamLength, sp,
                        tNbX, tNbY, stNbX, stNbY, selfStreamLength, lightSample, selfResInfo.mvy);
                else
                    misWeight = 1.0f / selfStreamLength;
                directCont += (misWeight * selfResInfo.mvx * selfStreamLength) * cont;
                combinedReservoir = selfRes;
                selectedTargetDensity = targetDensity;
                selectedMisWeight = misWeight;
                setSelected(sampleVis & SV_NEW);
            }
            combinedStreamLength = selfStreamLength;
        }

        auto mergeNeighbor = [&](int nbx, int nby, bool isTemporal) {
            const size_t nbPix = (size_t)nby * f->W + nbx;
            const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
            const GB1 neighborInfo = f->reservoirInfo[prevResIndex][nbPix];
            const uint32_t nbStreamLength = std::min(neighbor.streamLength, maxPrevStreamLength);
            if (neighborInfo.mvx > 0.0f) {
                const LightSample nbLightSample = neighbor.sample;
                const float3 cont = performDirectLighting<false>(s, sp.positionInWorld, sp.vOutLocal, sp.shadingFrame, sp.bsdf, nbLightSample);
                const float targetDensity = convertToWeight(cont);
                const float misWeight = isTemporal
                    ? computeMISWeight<SampleType::Temporal, withTemporalRIS, withSpatialRIS>(
                          f, p, prevCamera, prevBufIdx, prevResIndex, maxPrevStreamLength, sampleVis, selfStreamLength, sp,
                          tNbX, tNbY, stNbX, stNbY, nbStreamLength, nbLightSample, neighborInfo.mvy)
                    : computeMISWeight<SampleType::Spatiotemporal, withTemporalRIS, withSpatialRIS>(
                          f, p, prevCamera, prevBufIdx, prevResIndex, maxPrevStreamLength, sampleVis, selfStreamLength, sp,
                          tNbX, tNbY, stNbX, stNbY, nbStreamLength, nbLightSample, neighborInfo.mvy);
                const float weight = targetDensity * neighborInfo.mvx * nbStreamLength;
                const uint32_t visBit = isTemporal ? ((sampleVis & SV_T) ? 1u : 0u) : ((sampleVis & SV_ST) ? 1u : 0u);
                directCont += (visBit * misWeight * neighborInfo.mvx * nbStreamLength) * cont;
                if (combinedReservoir.update(nbLightSample, weight, rng.getFloat0cTo1o())) {
                    selectedTargetDensity = targetDensity;
                    selectedMisWeight = misWeight;
                    setSelected(visBit != 0);
                }
            }
            combinedStreamLength += nbStreamLength;
        };
        if (withTemporalRIS && (sampleVis & SV_T_PASSED))
            mergeNeighbor(tNbX, tNbY, true);
        if (withSpatialRIS && (sampleVis & SV_ST_PASSED))
            mergeNeighbor(stNbX, stNbY, false);

        combinedReservoir.streamLength = combinedStreamLength;
        contribution += directCont;

        float recPDFEstimate = selectedMisWeight * combinedReservoir.sumWeights / selectedTargetDensity;
        if (!std::isfinite(recPDFEstimate) || (p->reuseVisibility && !(sampleVis & SV_SELECTED))) {
            recPDFEstimate = 0.0f;
            selectedTargetDensity = 0.0f;
        }
        r->sampleVis[curBufIdx][pix] = sampleVis;
        storeReservoir(f, curResIndex, pix, combinedReservoir);
        f->reservoirInfo[curResIndex][pix] = GB1{ recPDFEstimate, selectedTargetDensity };
        f->rng[pix] = rng.state;
    }
    else if (useEnvLight(s, p)) { // optix_restir_di_rearch_kernels.cu:648-656
        const float2 texCoord = decodeTexCoords(gb3.qTexCoord);
        contribution = p->envLightPowerCoeff * s->env.fetch(texCoord.x, texCoord.y);
    }

    float3 prevColorResult(0.0f);
    if (p->numAccumFrames > 0)
        prevColorResult = float3(f->beauty[pix].x, f->beauty[pix].y, f->beauty[pix].z);
    const float curWeight = 1.0f / (1 + p->numAccumFrames);
    const float3 colorResult = (1 - curWeight) * prevColorResult + curWeight * contribution;
    f->beauty[pix] = F4{ colorResult.x, colorResult.y, colorResult.z, 1.0f };
}

// returns the number of shadow rays traced (GFX_RESTIR_TRACE_SHADOW_RAYS), 0 otherwise
static uint64_t restirRearch(orc_frame* f, const GfxFrameParams* p, int pass, int numThreads) {
    const Camera camera = makeCamera(p->camera);
    const Camera prevCamera = makeCamera(p->prevCamera);
    const uint32_t W = f->W, H = f->H;
    const bool T = p->enableTemporalReuse != 0, S = p->enableSpatialReuse != 0, U = p->useUnbiasedEstimator != 0;
    if (pass == GFX_RESTIR_PRESAMPLE_LIGHTS) {
        performLightPreSampling(f, p, numThreads);
        return 0;
    }
    rearchState(f);
    uint64_t rays = 0;
    if (pass == GFX_RESTIR_PER_PIXEL_RIS) {
        // every pixel reads the *incoming* RNG state of its tile origin, whatever that pixel stores afterwards
        const std::vector<uint64_t> rngIn = f->rng;
#pragma omp parallel for schedule(dynamic, 2) num_threads(numThreads)
        for (int64_t y = 0; y < (int64_t)H; ++y)
            for (uint32_t x = 0; x < W; ++x)
                performPerPixelRIS(f, p, camera, rngIn, x, (uint32_t)y);
        return 0;
    }
#pragma omp parallel for schedule(dynamic, 2) num_threads(numThreads) reduction(+ : rays)
    for (int64_t yy = 0; yy < (int64_t)H; ++yy) {
        const uint32_t y = (uint32_t)yy;
        uint64_t rowRays = 0;
        for (uint32_t x = 0; x < W; ++x) {
            if (pass == GFX_RESTIR_TRACE_SHADOW_RAYS) {
                if (!T && !S) traceShadowRays<false, false, false>(f, p, camera, prevCamera, x, y, &rowRays);
                else if (T && !S && !U) traceShadowRays<true, false, false>(f, p, camera, prevCamera, x, y, &rowRays);
                else if (!T && S && !U) traceShadowRays<false, true, false>(f, p, camera, prevCamera, x, y, &rowRays);
                else if (T && S && !U) traceShadowRays<true, true, false>(f, p, camera, prevCamera, x, y, &rowRays);
                else if (T && !S && U) traceShadowRays<true, false, true>(f, p, camera, prevCamera, x, y, &rowRays);
                else if (!T && S && U) traceShadowRays<false, true, true>(f, p, camera, prevCamera, x, y, &rowRays);
                else traceShadowRays<true, true, true>(f, p, camera, prevCamera, x, y, &rowRays);
            }
            else if (pass == GFX_RESTIR_SHADE_AND_RESAMPLE) {
                if (!T && !S) shadeAndResample<false, false>(f, p, camera, prevCamera, x, y);
                else if (T && !S) shadeAndResample<true, false>(f, p, camera, prevCamera, x, y);
                else if (!T && S) shadeAndResample<false, true>(f, p, camera, prevCamera, x, y);
                else shadeAndResample<true, true>(f, p, camera, prevCamera, x, y);
            }
        }
        rays += rowRays;
    }
    return rays;
}
