// restir.cu — ReSTIR DI passes as sm_100a kernels.
//
// Replaces restir.optixPipeline.launch(W, H, 1) with the ray-generation entry points
// performInitialRIS / performInitialAndTemporalRIS{Biased,Unbiased} / performSpatialRIS{Biased,
// Unbiased} / shading (restir_di/gpu_kernels/optix_restir_di_kernels.cu:14-287, 303-547, 559-637;
// launch sites restir_di/restir_di_main.cpp:2378-2421) and the helpers they share
// (sampleLight<false> restir_di_shared.h:320-516, performDirectLighting :518-557,
// evaluateVisibility :559-582, testNeighbor :747-771).  Visibility rays run through the inline
// any-hit traversal of traverse.cuh instead of optixTrace + the AH program.
//
// Per-pixel state is SoA: reservoirs are three float4 planes (Le|sumW, pos|M, n|-) so a pixel's own
// reservoir is three coalesced 16-byte loads; neighbour gathers hit the same planes through L2.
// RNG draw order per pixel is the reference's.
#include "restir_common.cuh"

namespace gfx {

// PHASE 0: the whole program in one kernel, visibility traced inline (megakernel form of the reference).
// PHASE 1: candidates + recPDF, the visibility ray is only *requested* (wavefront form).
// PHASE 2: applies the answered visibility, then the temporal merge.
//
// PHASE 1 runs the candidate loop through a per-warp survivor queue in shared memory.  ncu on the lock-step loop (round 2,
// profiles/r02_summary.md): ~70 % of the candidates are dark and leave after the staged light fetch, so the evaluation of the
// others (BSDF, geometry term, reservoir update: 58 % of the kernel's issue slots) ran with 9.4 of 32 lanes.  Here a lane only
// *generates* its pixel's candidate i (4 RNG draws, flattened light pick, staged fetch with the dark tests) and appends the
// survivors - light sample, density, the acceptance draw, the owning lane - to the warp's queue (ballot + popc, no atomics);
// whenever 32 are queued every lane evaluates one of them against the OWNER's shading state (also in shared memory), and the
// owners then apply their results in queue order, which is candidate order - so each pixel sees exactly the reference's
// sequence of Reservoir::update calls (restir_di_shared.h:118-125) and draws.
constexpr uint32_t kRisPixWords = 26, kRisEntWords = 16, kRisQueueSlots = 64;
struct RisWarpShared {
    float pix[kRisPixWords][32];             // owner state: position 0-2, vOutLocal 3-5, tangent 6-8, bitangent 9-11, normal 12-14,
                                             // bsdf: type 15, diffuse 16-18, specularF0 19-21, roughness 22
    float ent[kRisEntWords][kRisQueueSlots]; // 0 owner lane, 1 u, 2 ul, 3 u0, 4 u1, 5-7 position (-> weight, targetDensity, skip
                                             // after evaluation), 8-10 normal, 11-13 emittance, 14 density
};

struct RisSelection {
    float ul, u0, u1;
    bool has;
    float sumWeights, targetDensity;
};

// evaluates queue entries [0, n) (n <= 32), lets the owners apply them in order, moves the rest of the queue to the front
GFX_D void risFlush(const DevScene &s, RisWarpShared &w, uint32_t lane, uint32_t n, uint32_t* qCount, uint64_t* mySlots,
                    RisSelection* sel) {
    __syncwarp();
    if (lane < n) {
        const uint32_t owner = __float_as_uint(w.ent[0][lane]);
        LightSample ls;
        ls.position = f3(w.ent[5][lane], w.ent[6][lane], w.ent[7][lane]);
        ls.normal = f3(w.ent[8][lane], w.ent[9][lane], w.ent[10][lane]);
        ls.emittance = f3(w.ent[11][lane], w.ent[12][lane], w.ent[13][lane]);
        ls.atInfinity = 0;
        const float probDensity = w.ent[14][lane];
        const f3 positionInWorld(w.pix[0][owner], w.pix[1][owner], w.pix[2][owner]);
        const f3 vOutLocal(w.pix[3][owner], w.pix[4][owner], w.pix[5][owner]);
        ReferenceFrame shadingFrame;
        shadingFrame.tangent = f3(w.pix[6][owner], w.pix[7][owner], w.pix[8][owner]);
        shadingFrame.bitangent = f3(w.pix[9][owner], w.pix[10][owner], w.pix[11][owner]);
        shadingFrame.normal = f3(w.pix[12][owner], w.pix[13][owner], w.pix[14][owner]);
        BSDF bsdf;
        bsdf.type = __float_as_uint(w.pix[15][owner]);
        bsdf.diffuseColor = f3(w.pix[16][owner], w.pix[17][owner], w.pix[18][owner]);
        bsdf.specularF0Color = f3(w.pix[19][owner], w.pix[20][owner], w.pix[21][owner]);
        bsdf.roughness = w.pix[22][owner];
        // Dead candidates (light faces away, or lies below the shading horizon within the margin of the dark tests)
        // contribute exactly RGB(0) (performDirectLighting returns RGB(0) before any arithmetic when lpCos <= 0, and the
        // BRDFs return RGB(0) when vGiven.z * vSampled.z <= 0), so weight = +0 / probDensity = +0 and the reservoir is
        // untouched: skip the three IEEE divisions, which would all take the 0/x slow path.
        const f3 cont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, ls);
        const bool deadCont = cont.x == 0.0f && cont.y == 0.0f && cont.z == 0.0f;
        const bool skip = deadCont && probDensity > 0.0f;
        const float targetDensity = convertToWeight(cont);
        w.ent[5][lane] = skip ? 0.0f : targetDensity / probDensity;
        w.ent[6][lane] = targetDensity;
        w.ent[7][lane] = skip ? 1.0f : 0.0f;
    }
    __syncwarp();
    uint32_t mine = (uint32_t)(*mySlots & ((n >= 32 ? 0ull : (1ull << n)) - 1ull));
    while (mine) {
        const uint32_t slot = (uint32_t)__ffs((int)mine) - 1u;
        mine &= mine - 1u;
        if (w.ent[7][slot] != 0.0f)
            continue;
        const float weight = w.ent[5][slot];
        sel->sumWeights += weight;
        if (w.ent[1][slot] < weight / sel->sumWeights) {
            sel->ul = w.ent[2][slot];
            sel->u0 = w.ent[3][slot];
            sel->u1 = w.ent[4][slot];
            sel->has = true;
            sel->targetDensity = w.ent[6][slot];
        }
    }
    __syncwarp();
    const uint32_t rest = *qCount - n;
    if (rest) { // only after a full flush (n == 32): slots 32 .. 32 + rest - 1 move to the front
        float moved[kRisEntWords];
        if (lane < rest) {
#pragma unroll
            for (uint32_t k = 0; k < kRisEntWords; ++k)
                moved[k] = w.ent[k][n + lane];
        }
        __syncwarp();
        if (lane < rest) {
#pragma unroll
            for (uint32_t k = 0; k < kRisEntWords; ++k)
                w.ent[k][lane] = moved[k];
        }
    }
    *mySlots = n >= 32 ? (*mySlots >> 32) : 0ull;
    *qCount = rest;
}

template <bool withTemporalRIS, bool useUnbiasedEstimator, int PHASE, bool QUEUE>
GFX_D bool risPixel(const DevScene &s, const DevFrame &f, const DevFrameParams &p, RayRequest* request, RisWarpShared* warpShared) {
    // optix_restir_di_kernels.cu:14-287
    uint32_t x = blockIdx.x * 8 + threadIdx.x;
    uint32_t y = p.y0 + blockIdx.y * 8 + threadIdx.y;
    bool valid = x < f.W && y < p.y1;
    if (!QUEUE && !valid)
        return false;
    if (!valid) { // the queued form keeps the lane for the warp's queue: it reads a pixel of the image and writes nothing
        x = min(x, f.W - 1);
        y = min(y, p.y1 - 1);
    }
    const size_t pix = (size_t)y * f.W + x;
    const uint32_t curBufIdx = p.bufferIndex;

    const uint4 gb0 = f.gb0[curBufIdx][pix];
    valid = valid && gb0.x != 0xFFFFFFFFu;
    if (!QUEUE && !valid)
        return false;
    const float4 gb2 = f.gb2[curBufIdx][pix];
    const uint4 gb3 = f.gb3[curBufIdx][pix];

    f3 positionInWorld(gb2.x, gb2.y, gb2.z);
    const f3 geometricNormalInWorld = decodeVector(__float_as_uint(gb2.w));

    PCG32RNG rng{ f.rng[pix] };

    f3 vOut = p.camera.position - positionInWorld;
    const float frontHit = dot(vOut, geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
    positionInWorld = offsetRayOrigin(positionInWorld, frontHit * geometricNormalInWorld);
    const float dist = length(vOut);
    vOut /= dist;

    const f3 shadingNormalInWorld = decodeVector(gb3.x);
    const f3 shadingTangentInWorld = decodeVector(gb3.y);
    const ReferenceFrame shadingFrame(shadingNormalInWorld, shadingTangentInWorld);
    const f3 vOutLocal = shadingFrame.toLocal(vOut);
    const BSDF bsdf = setupBsdf(s, valid ? gb3.w : 0u);

    const uint32_t curResIndex = p.currentReservoirIndex;
    Reservoir reservoir;
    reservoir.initialize(emptyLightSample());

    float selectedTargetDensity = 0.0f;
    float recPDFEstimate = 0.0f;
    if (PHASE != 2) {
    const uint32_t numCandidates = 1u << p.log2NumCandidateSamples;
    // Streaming RIS (:66-123).  Reservoir::update (restir_di_shared.h:118-125) is spelled out: instead of
    // copying the 10-float LightSample on every acceptance, the three primary sample values that
    // generated it are kept and the winner is re-generated once after the loop (sampleLight is a pure
    // function of (ul, u0, u1)), which is bit-identical and frees registers in the hot loop.
    float selUl = 0.0f, selU0 = 0.0f, selU1 = 0.0f;
    bool hasSelection = false;
    float sumWeights = 0.0f;
    if constexpr (QUEUE) {
        RisWarpShared &w = *warpShared;
        const uint32_t lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31u;
        {
            const float state[kRisPixWords] = { positionInWorld.x, positionInWorld.y, positionInWorld.z, vOutLocal.x, vOutLocal.y, vOutLocal.z,
                shadingFrame.tangent.x, shadingFrame.tangent.y, shadingFrame.tangent.z,
                shadingFrame.bitangent.x, shadingFrame.bitangent.y, shadingFrame.bitangent.z,
                shadingFrame.normal.x, shadingFrame.normal.y, shadingFrame.normal.z, __uint_as_float(bsdf.type),
                bsdf.diffuseColor.x, bsdf.diffuseColor.y, bsdf.diffuseColor.z,
                bsdf.specularF0Color.x, bsdf.specularF0Color.y, bsdf.specularF0Color.z, bsdf.roughness, 0.0f, 0.0f, 0.0f };
#pragma unroll
            for (uint32_t k = 0; k < 23; ++k)
                w.pix[k][lane] = state[k];
        }
        RisSelection sel{ 0.0f, 0.0f, 0.0f, false, 0.0f, 0.0f };
        uint32_t qCount = 0;
        uint64_t mySlots = 0;
        for (uint32_t i = 0; i < numCandidates; ++i) {
            bool survivor = false;
            float ul = 0.0f, u0 = 0.0f, u1 = 0.0f, u = 0.0f, probDensity = 0.0f;
            LightSample lightSample = emptyLightSample();
            if (valid) {
                ul = rng.getFloat0cTo1o();
                u0 = rng.getFloat0cTo1o();
                u1 = rng.getFloat0cTo1o();
                // staged fetch of the light record: certainly-dark candidates stop early (their contribution is exactly
                // RGB(0): the reservoir is untouched and only the acceptance draw is consumed)
                survivor = !sampleLightUnlessDark(s, ul, u0, u1, positionInWorld, shadingNormalInWorld, vOutLocal.z, &lightSample, &probDensity);
                u = rng.getFloat0cTo1o();
            }
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, survivor);
            if (survivor) {
                const uint32_t slot = qCount + __popc(mask & ((1u << lane) - 1u));
                w.ent[0][slot] = __uint_as_float(lane);
                w.ent[1][slot] = u;
                w.ent[2][slot] = ul;
                w.ent[3][slot] = u0;
                w.ent[4][slot] = u1;
                w.ent[5][slot] = lightSample.position.x;
                w.ent[6][slot] = lightSample.position.y;
                w.ent[7][slot] = lightSample.position.z;
                w.ent[8][slot] = lightSample.normal.x;
                w.ent[9][slot] = lightSample.normal.y;
                w.ent[10][slot] = lightSample.normal.z;
                w.ent[11][slot] = lightSample.emittance.x;
                w.ent[12][slot] = lightSample.emittance.y;
                w.ent[13][slot] = lightSample.emittance.z;
                w.ent[14][slot] = probDensity; // probToSampleCurLightType = 1
                mySlots |= 1ull << slot;
            }
            qCount += __popc(mask);
            if (qCount >= 32)
                risFlush(s, w, lane, 32, &qCount, &mySlots, &sel);
        }
        if (qCount)
            risFlush(s, w, lane, qCount, &qCount, &mySlots, &sel);
        selUl = sel.ul;
        selU0 = sel.u0;
        selU1 = sel.u1;
        hasSelection = sel.has;
        sumWeights = sel.sumWeights;
        selectedTargetDensity = sel.targetDensity;
        if (!valid)
            return false;
    }
    else {
    for (uint32_t i = 0; i < numCandidates; ++i) {
        const float ul = rng.getFloat0cTo1o();
        const float probToSampleCurLightType = 1.0f;
        LightSample lightSample = emptyLightSample();
        float probDensity;
        const float u0 = rng.getFloat0cTo1o();
        const float u1 = rng.getFloat0cTo1o();
        // staged fetch of the light triangle: certainly-dark candidates stop early (their contribution is exactly
        // RGB(0): the reservoir is untouched and only the acceptance draw is consumed, as in the dead path below)
        if (sampleLightUnlessDark(s, ul, u0, u1, positionInWorld, shadingNormalInWorld, vOutLocal.z, &lightSample, &probDensity)) {
            (void)rng.getFloat0cTo1o();
            continue;
        }
        probDensity *= probToSampleCurLightType;
        // Dead candidates (light faces away, or lies below the shading horizon: ~70 % of them) contribute exactly
        // RGB(0) (performDirectLighting returns RGB(0) before any arithmetic when lpCos <= 0, and the BRDFs
        // return RGB(0) when vGiven.z * vSampled.z <= 0), so weight = +0 / probDensity = +0 and the reservoir
        // is untouched: skip the three IEEE divisions, which would all take the 0/x slow path.
        const f3 cont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, lightSample);
        const bool deadCont = cont.x == 0.0f && cont.y == 0.0f && cont.z == 0.0f;
        if (deadCont && probDensity > 0.0f) {
            (void)rng.getFloat0cTo1o();
            continue;
        }
        const float targetDensity = convertToWeight(cont);
        const float weight = targetDensity / probDensity;
        const float u = rng.getFloat0cTo1o();
        sumWeights += weight;
        if (u < weight / sumWeights) {
            selUl = ul;
            selU0 = u0;
            selU1 = u1;
            hasSelection = true;
            selectedTargetDensity = targetDensity;
        }
    }
    }
    if (hasSelection) {
        float unusedDensity;
        sampleLight(s, selUl, selU0, selU1, &reservoir.sample, &unusedDensity);
    }
    reservoir.sumWeights = sumWeights;
    reservoir.streamLength = numCandidates;

    recPDFEstimate = reservoir.sumWeights / (selectedTargetDensity * reservoir.streamLength);
    if (!isfinite(recPDFEstimate)) {
        recPDFEstimate = 0.0f;
        selectedTargetDensity = 0.0f;
    }
    }
    else { // PHASE 2: pick up where phase 1 stopped
        reservoir = loadReservoir(f, curResIndex, pix);
        const float2 info = f.reservoirInfo[curResIndex][pix];
        recPDFEstimate = info.x;
        selectedTargetDensity = info.y;
    }

    if (p.reuseVisibility && selectedTargetDensity > 0.0f) {
        if (PHASE == 1) {
            visibilityRay(positionInWorld, reservoir.sample, (uint32_t)pix, request);
            f.rng[pix] = rng.state;
            storeReservoir(f, curResIndex, pix, reservoir);
            f.reservoirInfo[curResIndex][pix] = make_float2(recPDFEstimate, selectedTargetDensity);
            return true;
        }
        const bool visible = PHASE == 2 ? (f.visibility[pix] != 0) : evaluateVisibility(s, positionInWorld, reservoir.sample);
        if (!visible) {
            recPDFEstimate = 0.0f;
            selectedTargetDensity = 0.0f;
        }
    }
    if (PHASE == 1) {
        f.rng[pix] = rng.state;
        storeReservoir(f, curResIndex, pix, reservoir);
        f.reservoirInfo[curResIndex][pix] = make_float2(recPDFEstimate, selectedTargetDensity);
        return false;
    }

    if (withTemporalRIS) {
        const uint32_t prevBufIdx = (curBufIdx + 1) % 2;
        const uint32_t prevResIndex = (curResIndex + 1) % 2;

        bool neighborIsSelected = false;
        const uint32_t selfStreamLength = reservoir.streamLength;
        if (recPDFEstimate == 0.0f)
            reservoir.initialize(emptyLightSample());
        uint32_t combinedStreamLength = selfStreamLength;
        const uint32_t maxPrevStreamLength = 20 * selfStreamLength;

        const float2 gb1 = f.gb1[curBufIdx][pix];
        const int nbx = dm_f2int(x + 0.5f - gb1.x);
        const int nby = dm_f2int(y + 0.5f - gb1.y);

        const bool acceptedNeighbor = testNeighbor<!useUnbiasedEstimator>(f, p.camera, prevBufIdx, nbx, nby, dist, shadingNormalInWorld);
        if (acceptedNeighbor) {
            const size_t nbPix = (size_t)nby * f.W + nbx;
            const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
            const float2 neighborInfo = f.reservoirInfo[prevResIndex][nbPix];
            const LightSample nbLightSample = neighbor.sample;
            const f3 cont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, nbLightSample);
            const float targetDensity = convertToWeight(cont);
            const uint32_t nbStreamLength = min(neighbor.streamLength, maxPrevStreamLength);
            const float weight = targetDensity * neighborInfo.x * nbStreamLength;
            if (reservoir.update(nbLightSample, weight, rng.getFloat0cTo1o())) {
                selectedTargetDensity = targetDensity;
                if (useUnbiasedEstimator)
                    neighborIsSelected = true;
            }
            combinedStreamLength += nbStreamLength;
        }
        reservoir.streamLength = combinedStreamLength;

        float weightForEstimate;
        if (useUnbiasedEstimator) { // :192-266, useMIS_RIS = true
            const LightSample selectedLightSample = reservoir.sample;
            float numWeight, denomWeight;
            {
                const f3 cont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, selectedLightSample);
                const float targetDensityForSelf = convertToWeight(cont);
                numWeight = targetDensityForSelf;
                denomWeight = targetDensityForSelf * selfStreamLength;
            }
            if (acceptedNeighbor) {
                const size_t nbPix = (size_t)nby * f.W + nbx;
                const float4 nbGb2 = f.gb2[prevBufIdx][nbPix];
                const uint4 nbGb3 = f.gb3[prevBufIdx][nbPix];
                f3 nbPositionInWorld(nbGb2.x, nbGb2.y, nbGb2.z);
                const f3 nbGeometricNormalInWorld = decodeVector(__float_as_uint(nbGb2.w));
                const f3 nbVOut = normalize(p.prevCamera.position - nbPositionInWorld);
                const float nbFrontHit = dot(nbVOut, nbGeometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
                nbPositionInWorld = offsetRayOrigin(nbPositionInWorld, nbFrontHit * nbGeometricNormalInWorld);
                const BSDF nbBsdf = setupBsdf(s, nbGb3.w);
                const ReferenceFrame nbShadingFrame(decodeVector(nbGb3.x), decodeVector(nbGb3.y));
                const f3 nbVOutLocal = nbShadingFrame.toLocal(nbVOut);
                const Reservoir neighbor = loadReservoir(f, prevResIndex, nbPix);
                const f3 cont = performDirectLighting<false>(s, nbPositionInWorld, nbVOutLocal, nbShadingFrame, nbBsdf, selectedLightSample);
                const float nbTargetDensity = convertToWeight(cont);
                const uint32_t nbStreamLength = min(neighbor.streamLength, maxPrevStreamLength);
                denomWeight += nbTargetDensity * nbStreamLength;
                if (neighborIsSelected)
                    numWeight = nbTargetDensity;
            }
            weightForEstimate = numWeight / denomWeight;
        }
        else {
            weightForEstimate = 1.0f / reservoir.streamLength;
        }

        recPDFEstimate = weightForEstimate * reservoir.sumWeights / selectedTargetDensity;
        if (!isfinite(recPDFEstimate)) {
            recPDFEstimate = 0.0f;
            selectedTargetDensity = 0.0f;
        }
    }

    f.rng[pix] = rng.state;
    storeReservoir(f, curResIndex, pix, reservoir);
    f.reservoirInfo[curResIndex][pix] = make_float2(recPDFEstimate, selectedTargetDensity);
    return false;
}

template <bool withTemporalRIS, bool useUnbiasedEstimator, int PHASE, bool QUEUE = false>
__global__ void __launch_bounds__(64) k_initialAndTemporalRIS(DevScene s, DevFrame f, DevFrameParams p) {
    static_assert(!QUEUE || PHASE == 1, "the survivor queue belongs to the wavefront candidate pass");
    RayRequest request;
    RisWarpShared* warpShared = nullptr;
    if constexpr (QUEUE) {
        __shared__ RisWarpShared shared[2]; // one per warp of the 8x8 block (rows 0-3, rows 4-7)
        warpShared = &shared[threadIdx.y >> 2];
    }
    const bool want = risPixel<withTemporalRIS, useUnbiasedEstimator, PHASE, QUEUE>(s, f, p, &request, warpShared);
    if (PHASE == 1)
        enqueueRay(f, s.rayCounter, want, request);
}

// ---------------------------------------------------------------------------------------------
GFX_D void neighborCoord(const DevFrame &f, const DevFrameParams &p, uint32_t x, uint32_t y, int nIdx, PCG32RNG &rng,
                         int* nbx, int* nby) { // optix_restir_di_kernels.cu:356-371
    float radius = p.spatialNeighborRadius;
    float deltaX, deltaY;
    if (p.useLowDiscrepancyNeighbors) {
        const float2 delta = __ldg(f.neighborDeltas + ((p.spatialNeighborBaseIndex + nIdx) % 1024));
        deltaX = radius * delta.x;
        deltaY = radius * delta.y;
    }
    else {
        radius *= sqrtf(rng.getFloat0cTo1o());
        const float angle = 2 * kPi * rng.getFloat0cTo1o();
        float sa, ca;
        dm_sincos(angle, &sa, &ca);
        deltaX = radius * ca;
        deltaY = radius * sa;
    }
    *nbx = dm_f2int(x + 0.5f + deltaX);
    *nby = dm_f2int(y + 0.5f + deltaY);
}

template <bool useUnbiasedEstimator>
__global__ void __launch_bounds__(64) k_spatialRIS(DevScene s, DevFrame f, DevFrameParams p) {
    // optix_restir_di_kernels.cu:303-547
    const uint32_t x = blockIdx.x * 8 + threadIdx.x;
    const uint32_t y = p.y0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= f.W || y >= p.y1)
        return;
    const size_t pix = (size_t)y * f.W + x;
    const uint32_t bufIdx = p.bufferIndex;

    const uint4 gb0 = f.gb0[bufIdx][pix];
    if (gb0.x == 0xFFFFFFFFu)
        return;
    const float4 gb2 = f.gb2[bufIdx][pix];
    const uint4 gb3 = f.gb3[bufIdx][pix];

    f3 positionInWorld(gb2.x, gb2.y, gb2.z);
    const f3 geometricNormalInWorld = decodeVector(__float_as_uint(gb2.w));
    PCG32RNG rng{ f.rng[pix] };

    f3 vOut = p.camera.position - positionInWorld;
    const float frontHit = dot(vOut, geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
    positionInWorld = offsetRayOrigin(positionInWorld, frontHit * geometricNormalInWorld);
    const float dist = length(vOut);
    vOut /= dist;

    const ReferenceFrame shadingFrame(decodeVector(gb3.x), decodeVector(gb3.y));
    const f3 vOutLocal = shadingFrame.toLocal(vOut);
    const BSDF bsdf = setupBsdf(s, gb3.w);

    const uint32_t srcResIndex = p.currentReservoirIndex;
    const uint32_t dstResIndex = (srcResIndex + 1) % 2;

    Reservoir combinedReservoir;
    combinedReservoir.initialize(emptyLightSample());
    float selectedTargetDensity = 0.0f;
    int32_t selectedNeighborIndex = -1;

    const Reservoir self = loadReservoir(f, srcResIndex, pix);
    const float2 selfResInfo = f.reservoirInfo[srcResIndex][pix];
    if (selfResInfo.x > 0.0f) {
        combinedReservoir = self;
        selectedTargetDensity = selfResInfo.y;
    }
    uint32_t combinedStreamLength = self.streamLength;

    for (int nIdx = 0; nIdx < (int)p.numSpatialNeighbors; ++nIdx) {
        int nbx, nby;
        neighborCoord(f, p, x, y, nIdx, rng, &nbx, &nby);
        const bool acceptedNeighbor =
            testNeighbor<!useUnbiasedEstimator>(f, p.camera, bufIdx, nbx, nby, dist, shadingFrame.normal)
            && (nbx != (int)x || nby != (int)y);
        if (acceptedNeighbor) {
            const size_t nbPix = (size_t)nby * f.W + nbx;
            const Reservoir neighbor = loadReservoir(f, srcResIndex, nbPix);
            const float2 neighborInfo = f.reservoirInfo[srcResIndex][nbPix];
            const LightSample nbLightSample = neighbor.sample;
            const f3 cont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, nbLightSample);
            const float targetDensity = convertToWeight(cont);
            const uint32_t nbStreamLength = neighbor.streamLength;
            const float weight = targetDensity * neighborInfo.x * nbStreamLength;
            if (combinedReservoir.update(nbLightSample, weight, rng.getFloat0cTo1o())) {
                selectedTargetDensity = targetDensity;
                if (useUnbiasedEstimator)
                    selectedNeighborIndex = nIdx;
            }
            combinedStreamLength += nbStreamLength;
        }
    }
    combinedReservoir.streamLength = combinedStreamLength;

    float weightForEstimate = 0.0f;
    if (useUnbiasedEstimator) { // :414-529
        if (selectedTargetDensity > 0.0f) {
            const LightSample selectedLightSample = combinedReservoir.sample;
            float numWeight, denomWeight;
            bool visibility = true;
            {
                f3 cont;
                if (p.reuseVisibility)
                    cont = performDirectLighting<true>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, selectedLightSample);
                else
                    cont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, selectedLightSample);
                const float targetDensityForSelf = convertToWeight(cont);
                if (p.reuseVisibility)
                    visibility = targetDensityForSelf > 0.0f;
                numWeight = targetDensityForSelf;
                denomWeight = targetDensityForSelf * self.streamLength;
            }
            for (int nIdx = 0; nIdx < (int)p.numSpatialNeighbors; ++nIdx) {
                int nbx, nby;
                neighborCoord(f, p, x, y, nIdx, rng, &nbx, &nby);
                const bool acceptedNeighbor =
                    (nbx >= 0 && nbx < (int)f.W && nby >= 0 && nby < (int)f.H) && (nbx != (int)x || nby != (int)y);
                if (acceptedNeighbor) {
                    const size_t nbPix = (size_t)nby * f.W + nbx;
                    if (f.gb0[bufIdx][nbPix].x == 0xFFFFFFFFu)
                        continue;
                    const float4 nbGb2 = f.gb2[bufIdx][nbPix];
                    const uint4 nbGb3 = f.gb3[bufIdx][nbPix];
                    f3 nbPositionInWorld(nbGb2.x, nbGb2.y, nbGb2.z);
                    const f3 nbGeometricNormalInWorld = decodeVector(__float_as_uint(nbGb2.w));
                    const f3 nbVOut = normalize(p.prevCamera.position - nbPositionInWorld);
                    const float nbFrontHit = dot(nbVOut, nbGeometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
                    nbPositionInWorld = offsetRayOrigin(nbPositionInWorld, nbFrontHit * nbGeometricNormalInWorld);
                    const BSDF nbBsdf = setupBsdf(s, nbGb3.w);
                    const ReferenceFrame nbShadingFrame(decodeVector(nbGb3.x), decodeVector(nbGb3.y));
                    const f3 nbVOutLocal = nbShadingFrame.toLocal(nbVOut);
                    const Reservoir neighbor = loadReservoir(f, srcResIndex, nbPix);
                    f3 cont;
                    if (p.reuseVisibility)
                        cont = performDirectLighting<true>(s, nbPositionInWorld, nbVOutLocal, nbShadingFrame, nbBsdf, selectedLightSample);
                    else
                        cont = performDirectLighting<false>(s, nbPositionInWorld, nbVOutLocal, nbShadingFrame, nbBsdf, selectedLightSample);
                    const float nbTargetDensity = convertToWeight(cont);
                    const uint32_t nbStreamLength = neighbor.streamLength;
                    denomWeight += nbTargetDensity * nbStreamLength;
                    if (nIdx == selectedNeighborIndex)
                        numWeight = nbTargetDensity;
                }
            }
            weightForEstimate = numWeight / denomWeight;
            if (p.reuseVisibility && !visibility)
                weightForEstimate = 0.0f;
        }
    }
    else {
        weightForEstimate = 1.0f / combinedReservoir.streamLength;
    }

    float recPDFEstimate = weightForEstimate * combinedReservoir.sumWeights / selectedTargetDensity;
    float targetDensityOut = selectedTargetDensity;
    if (!isfinite(recPDFEstimate)) {
        recPDFEstimate = 0.0f;
        targetDensityOut = 0.0f;
    }

    f.rng[pix] = rng.state;
    storeReservoir(f, dstResIndex, pix, combinedReservoir);
    f.reservoirInfo[dstResIndex][pix] = make_float2(recPDFEstimate, targetDensityOut);
}

// ---------------------------------------------------------------------------------------------
// wavefront form of `shading`: phase 1 only requests the shadow ray of the surviving sample
__global__ void __launch_bounds__(64) k_shadingRays(DevScene s, DevFrame f, DevFrameParams p) {
    RayRequest request;
    bool want = false;
    const uint32_t x = blockIdx.x * 8 + threadIdx.x;
    const uint32_t y = p.y0 + blockIdx.y * 8 + threadIdx.y;
    if (x < f.W && y < p.y1) {
        const size_t pix = (size_t)y * f.W + x;
        const uint32_t bufIdx = p.bufferIndex;
        if (f.gb0[bufIdx][pix].x != 0xFFFFFFFFu) {
            const float2 reservoirInfo = f.reservoirInfo[p.currentReservoirIndex][pix];
            const float recPDFEstimate = reservoirInfo.x;
            if (recPDFEstimate > 0 && isfinite(recPDFEstimate)) {
                const float4 gb2 = f.gb2[bufIdx][pix];
                f3 positionInWorld(gb2.x, gb2.y, gb2.z);
                const f3 geometricNormalInWorld = decodeVector(__float_as_uint(gb2.w));
                const f3 vOut = normalize(p.camera.position - positionInWorld);
                const float frontHit = dot(vOut, geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
                positionInWorld = offsetRayOrigin(positionInWorld, frontHit * geometricNormalInWorld);
                const Reservoir reservoir = loadReservoir(f, p.currentReservoirIndex, pix);
                visibilityRay(positionInWorld, reservoir.sample, (uint32_t)pix, &request);
                want = true;
            }
        }
    }
    enqueueRay(f, s.rayCounter, want, request);
}

template <bool QUEUED_VISIBILITY>
__global__ void __launch_bounds__(64) k_shading(DevScene s, DevFrame f, DevFrameParams p) {
    // optix_restir_di_kernels.cu:559-637
    const uint32_t x = blockIdx.x * 8 + threadIdx.x;
    const uint32_t y = p.y0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= f.W || y >= p.y1)
        return;
    const size_t pix = (size_t)y * f.W + x;
    const uint32_t bufIdx = p.bufferIndex;
    const uint4 gb0 = f.gb0[bufIdx][pix];
    const uint4 gb3 = f.gb3[bufIdx][pix];

    f3 contribution(0.01f, 0.01f, 0.01f);
    if (gb0.x != 0xFFFFFFFFu) {
        const float4 gb2 = f.gb2[bufIdx][pix];
        f3 positionInWorld(gb2.x, gb2.y, gb2.z);
        const f3 geometricNormalInWorld = decodeVector(__float_as_uint(gb2.w));
        const f3 vOut = normalize(p.camera.position - positionInWorld);
        const float frontHit = dot(vOut, geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
        positionInWorld = offsetRayOrigin(positionInWorld, frontHit * geometricNormalInWorld);

        const ReferenceFrame shadingFrame(decodeVector(gb3.x), decodeVector(gb3.y));
        const f3 vOutLocal = shadingFrame.toLocal(vOut);
        const GfxMaterialDesc* mat = s.materials + gb3.w;
        const BSDF bsdf = setupBsdf(s, gb3.w);

        const uint32_t curResIndex = p.currentReservoirIndex;
        const Reservoir reservoir = loadReservoir(f, curResIndex, pix);
        const float2 reservoirInfo = f.reservoirInfo[curResIndex][pix];

        contribution = f3(0.0f);
        if (vOutLocal.z > 0) {
            f3 emittance(0.0f);
            if (mat->hasEmittance)
                emittance = f3(mat->emittance[0], mat->emittance[1], mat->emittance[2]);
            contribution += emittance / kPi;
        }

        const LightSample lightSample = reservoir.sample;
        f3 directCont(0.0f);
        const float recPDFEstimate = reservoirInfo.x;
        if (recPDFEstimate > 0 && isfinite(recPDFEstimate)) {
            const bool visDone = p.reuseVisibility &&
                (!p.enableTemporalReuse || (p.enableSpatialReuse && p.useUnbiasedEstimator));
            if (visDone)
                directCont = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, lightSample);
            else if (QUEUED_VISIBILITY) // performDirectLighting<true> with the answer of the wavefront trace
                directCont = f.visibility[pix] ? performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, lightSample)
                                               : f3(0.0f);
            else
                directCont = performDirectLighting<true>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, lightSample);
        }
        contribution += recPDFEstimate * directCont;
    }

    f3 prevColorResult(0.0f);
    if (p.numAccumFrames > 0) {
        const float4 pb = f.beauty[pix];
        prevColorResult = f3(pb.x, pb.y, pb.z);
    }
    const float curWeight = 1.0f / (1 + p.numAccumFrames);
    const f3 colorResult = (1 - curWeight) * prevColorResult + curWeight * contribution;
    f.beauty[pix] = make_float4(colorResult.x, colorResult.y, colorResult.z, 1.0f);
}

int launchReSTIR(gfx_ctx* ctx, cudaStream_t stream, const GfxFrameParams* params, int pass) {
    const DevFrameParams p = makeDevParams(ctx, params);
    if (p.y1 <= p.y0)
        return GFX_OK;
    if (p.log2NumCandidateSamples > 15 || p.numSpatialNeighbors > 15)
        return GFX_ERR_INVALID_ARGUMENT; // 4-bit fields in the reference (restir_di_shared.h:258-259)
    const dim3 block(8, 8);
    const dim3 grid((ctx->frame.W + 7) / 8, (p.y1 - p.y0 + 7) / 8);
    const DevScene s = ctx->devScene();
    const DevFrame f = ctx->devFrame();
    // wavefront form (default): request rays -> persistent-thread trace -> resolve.  GFX_MEGAKERNEL=1 selects
    // the single-kernel form with inline traversal (same results; kept for A/B measurements).
    static const bool megakernel = getenv("GFX_MEGAKERNEL") != nullptr;
    const bool wave = !megakernel;
    // GFX_RIS_QUEUE=1: candidate loop through the per-warp survivor queue (see risFlush); 0 = lock-step loop
    static const bool risQueue = getenv("GFX_RIS_QUEUE") && getenv("GFX_RIS_QUEUE")[0] == '1';
#define RIS_LAUNCH(T, U) \
    if (wave && p.reuseVisibility) { \
        int rc_ = resetVisibilityQueue(ctx, stream); if (rc_) return rc_; \
        { GFX_TIMED(ctx, stream, "ris_candidates"); \
          if (risQueue) k_initialAndTemporalRIS<T, U, 1, true><<<grid, block, 0, stream>>>(s, f, p); \
          else k_initialAndTemporalRIS<T, U, 1, false><<<grid, block, 0, stream>>>(s, f, p); } ctx->launches++; \
        rc_ = traceVisibilityQueue(ctx, stream); if (rc_) return rc_; \
        { GFX_TIMED(ctx, stream, "ris_resolve_temporal"); k_initialAndTemporalRIS<T, U, 2><<<grid, block, 0, stream>>>(s, f, p); } \
    } else { \
        { GFX_TIMED(ctx, stream, "ris_megakernel"); k_initialAndTemporalRIS<T, U, 0><<<grid, block, 0, stream>>>(s, f, p); } \
    }
    switch (pass) {
    case GFX_RESTIR_INITIAL_RIS: RIS_LAUNCH(false, false); break;
    case GFX_RESTIR_INITIAL_AND_TEMPORAL_BIASED: RIS_LAUNCH(true, false); break;
    case GFX_RESTIR_INITIAL_AND_TEMPORAL_UNBIASED: RIS_LAUNCH(true, true); break;
    case GFX_RESTIR_SPATIAL_BIASED: { GFX_TIMED(ctx, stream, "spatial_ris"); k_spatialRIS<false><<<grid, block, 0, stream>>>(s, f, p); } break;
    case GFX_RESTIR_SPATIAL_UNBIASED: { GFX_TIMED(ctx, stream, "spatial_ris_unbiased"); k_spatialRIS<true><<<grid, block, 0, stream>>>(s, f, p); } break;
    case GFX_RESTIR_SHADING: {
        const bool visDone = p.reuseVisibility && (!p.enableTemporalReuse || (p.enableSpatialReuse && p.useUnbiasedEstimator));
        if (wave && !visDone) {
            int rc_ = resetVisibilityQueue(ctx, stream); if (rc_) return rc_;
            { GFX_TIMED(ctx, stream, "shading_rays"); k_shadingRays<<<grid, block, 0, stream>>>(s, f, p); } ctx->launches++;
            rc_ = traceVisibilityQueue(ctx, stream); if (rc_) return rc_;
            { GFX_TIMED(ctx, stream, "shading"); k_shading<true><<<grid, block, 0, stream>>>(s, f, p); }
        }
        else {
            { GFX_TIMED(ctx, stream, "shading_megakernel"); k_shading<false><<<grid, block, 0, stream>>>(s, f, p); }
        }
        break;
    }
    default:
        ctx->setError("gfx_restir_launch: unknown pass");
        return GFX_ERR_INVALID_ARGUMENT;
    }
#undef RIS_LAUNCH
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
