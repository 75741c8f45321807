// oracle/regir.inl — TEST INFRASTRUCTURE (CPU oracle), included by render.cpp after pathtrace.inl.
//
// CPU restatement of ReGIR (grid-based reservoirs for light sampling):
//   sampleIntensity                          regir/gpu_kernels/build_cell_reservoirs.cu:6-69
//   buildCellReservoirsAndTemporalReuse<>    regir/gpu_kernels/build_cell_reservoirs.cu:71-225
//   updateLastAccessFrameIndices             regir/gpu_kernels/build_cell_reservoirs.cu:235-248
//   calcCellLinearIndex                      regir/regir_shared.h:731-742
//   sampleFromCell / performNextEventEstimation<true>   regir/gpu_kernels/optix_pathtracing_kernels.cu:18-101
//   pathTrace_rayGen_generic<true> / pathTrace_closestHit_generic<true>   :154-307, 309-383
//   grid + slot RNG initialisation           regir/regir_main.cpp:1073-1109 (32x8x32 cells over initialSceneAabb,
//                                            512 slots per cell, mt19937_64(591842031321323413))
//
// Two spots of the reference are reproduced as written rather than as probably intended:
//  * sampleIntensity compares `lpCos` (still 1 at that point) with minSquaredDistance (:54-55), so the cosine is
//    evaluated whenever the squared half diagonal of a cell is below 1 and assumed to be 1 otherwise;
//  * the ReGIR closest-hit program weights implicit light hits by MIS against `hypAreaPDensity`, which
//    computeSurfacePoint<false, ...> never writes (:322-329, 343-356): an uninitialised read.  Oracle and
//    product define that value as 0, i.e. the implicit hit is taken with weight bsdfP^2 / (bsdfP^2 + 0).

static constexpr uint32_t kNumLightSlotsPerCell = 512; // regir_shared.h:7

struct RegirSlot { // Reservoir<LightSample> (48 B) + ReservoirInfo (8 B), padded to one 64-byte record
    float emittance[3], sumWeights;
    float position[3]; uint32_t streamLengthAtInf;
    float normal[3], recPDFEstimate;
    float targetDensity, pad[3];
};
static_assert(sizeof(RegirSlot) == 64, "RegirSlot is one 64-byte record");

struct orc_regir {
    uint32_t dim[3];
    uint32_t numCells, numSlots;
    float3 gridOrigin, gridCellSize;
    std::vector<RegirSlot> slots[2];
    std::vector<uint64_t> slotRngs;
    std::vector<uint32_t> perCellNumAccesses, lastAccessFrameIndices;
    uint32_t numActiveCells[2];
};

static void regirDestroy(orc_regir* r) { delete r; }

static orc_regir* regirState(orc_frame* f, const GfxFrameParams* p) {
    uint32_t dim[3] = { p->regirGridDim[0], p->regirGridDim[1], p->regirGridDim[2] };
    if (dim[0] == 0 || dim[1] == 0 || dim[2] == 0) {
        dim[0] = 32; dim[1] = 8; dim[2] = 32; // regir_main.cpp:1109
    }
    if (f->regir && f->regir->dim[0] == dim[0] && f->regir->dim[1] == dim[1] && f->regir->dim[2] == dim[2])
        return f->regir;
    if (f->regir)
        regirDestroy(f->regir);
    orc_regir* r = new orc_regir();
    std::memcpy(r->dim, dim, sizeof(dim));
    r->numCells = dim[0] * dim[1] * dim[2];
    r->numSlots = r->numCells * kNumLightSlotsPerCell;
    const float3 minP(p->sceneAabbMin[0], p->sceneAabbMin[1], p->sceneAabbMin[2]);
    const float3 maxP(p->sceneAabbMax[0], p->sceneAabbMax[1], p->sceneAabbMax[2]);
    r->gridOrigin = minP;
    r->gridCellSize = (maxP - minP) / float3((float)dim[0], (float)dim[1], (float)dim[2]);
    RegirSlot empty;
    std::memset(&empty, 0, sizeof(empty));
    for (int i = 0; i < 2; ++i)
        r->slots[i].assign(r->numSlots, empty);
    r->slotRngs.resize(r->numSlots);
    std::mt19937_64 rngSeed(591842031321323413ull);
    for (auto &s : r->slotRngs)
        s = rngSeed();
    r->perCellNumAccesses.assign(r->numCells, 0u);
    r->lastAccessFrameIndices.assign(r->numCells, 0xFFFFFFFFu); // fill(frameIndex = -1)
    r->numActiveCells[0] = r->numActiveCells[1] = 0;
    f->regir = r;
    return r;
}

static inline Reservoir loadSlot(const RegirSlot &s) {
    Reservoir r;
    r.sample.emittance = float3(s.emittance[0], s.emittance[1], s.emittance[2]);
    r.sample.position = float3(s.position[0], s.position[1], s.position[2]);
    r.sample.normal = float3(s.normal[0], s.normal[1], s.normal[2]);
    r.sample.atInfinity = s.streamLengthAtInf >> 31;
    r.sumWeights = s.sumWeights;
    r.streamLength = s.streamLengthAtInf & 0x7FFFFFFFu;
    return r;
}
static inline void storeSlot(RegirSlot &s, const Reservoir &r, float recPDFEstimate, float targetDensity) {
    s.emittance[0] = r.sample.emittance.x; s.emittance[1] = r.sample.emittance.y; s.emittance[2] = r.sample.emittance.z;
    s.sumWeights = r.sumWeights;
    s.position[0] = r.sample.position.x; s.position[1] = r.sample.position.y; s.position[2] = r.sample.position.z;
    s.streamLengthAtInf = (r.streamLength & 0x7FFFFFFFu) | (r.sample.atInfinity << 31);
    s.normal[0] = r.sample.normal.x; s.normal[1] = r.sample.normal.y; s.normal[2] = r.sample.normal.z;
    s.recPDFEstimate = recPDFEstimate;
    s.targetDensity = targetDensity;
    s.pad[0] = s.pad[1] = s.pad[2] = 0.0f;
}
static inline LightSample emptyLightSample() { // LightSample(): atInfinity(false); the other members are zeroed here
    LightSample ls;
    ls.emittance = float3(0.0f);
    ls.position = float3(0.0f);
    ls.normal = float3(0.0f);
    ls.atInfinity = 0;
    return ls;
}

// build_cell_reservoirs.cu:6-69
static float3 sampleIntensity(const orc_scene* s, const GfxFrameParams* p, const float3 &cellCenter, const float3 &halfCellSize,
                              float minSquaredDistance, float uLight, bool sampleEnvLight, float uPos0, float uPos1,
                              LightSample* lightSample, float* probDensity) {
    sampleLight(s, p, uLight, sampleEnvLight, uPos0, uPos1, lightSample, probDensity);
    float dist2 = minSquaredDistance;
    float lpCos = 1;
    const bool isOutsideCell =
        lightSample->atInfinity ||
        lightSample->position.x < cellCenter.x - halfCellSize.x ||
        lightSample->position.x > cellCenter.x + halfCellSize.x ||
        lightSample->position.y < cellCenter.y - halfCellSize.y ||
        lightSample->position.y > cellCenter.y + halfCellSize.y ||
        lightSample->position.z < cellCenter.z - halfCellSize.z ||
        lightSample->position.z > cellCenter.z + halfCellSize.z;
    if (isOutsideCell) {
        const float3 shadowRayDir = lightSample->atInfinity ? lightSample->position : (lightSample->position - cellCenter);
        const float perpDistance = dot(-shadowRayDir, lightSample->normal);
        dist2 = sqLength(shadowRayDir);
        const float dist = std::sqrt(dist2);
        const bool cellIsInValidHalfSpace = lpCos > minSquaredDistance || lightSample->atInfinity;
        const bool cellIsInInvalidHalfSpace = lpCos < -minSquaredDistance;
        if (cellIsInValidHalfSpace)
            lpCos = perpDistance / dist;
        else if (cellIsInInvalidHalfSpace)
            lpCos = 0.0f;
    }
    if (lpCos > 0.0f) {
        const float3 Le = lightSample->emittance / kPi;
        return Le * (lpCos / dist2);
    }
    return float3(0.0f);
}

// build_cell_reservoirs.cu:71-233
extern "C" void orc_regir_build_cells(orc_frame* f, const GfxFrameParams* p, uint32_t frameIndex, int useTemporalReuse, int numThreads) {
    if (numThreads <= 0) numThreads = omp_get_max_threads();
    const orc_scene* s = f->scene;
    orc_regir* r = regirState(f, p);
    const uint32_t bufferIndex = p->bufferIndex & 1, prevBufferIndex = (bufferIndex + 1) % 2;
    r->numActiveCells[bufferIndex] = 0;
    const float3 halfCellSize = 0.5f * r->gridCellSize;
    const float minSquaredDistance = sqLength(0.5f * r->gridCellSize);
    const uint32_t numCandidates = 1u << p->regirLog2NumCandidatesPerLightSlot;

#pragma omp parallel for schedule(dynamic, 1) num_threads(numThreads)
    for (int64_t cell = 0; cell < (int64_t)r->numCells; ++cell) {
        const uint32_t cellLinearIndex = (uint32_t)cell;
        const uint32_t lastAccessFrameIndex = r->lastAccessFrameIndices[cellLinearIndex];
        r->perCellNumAccesses[cellLinearIndex] = 0;
        if (frameIndex - lastAccessFrameIndex > 8)
            continue;
        const uint32_t iz = cellLinearIndex / (r->dim[0] * r->dim[1]);
        const uint32_t iy = (cellLinearIndex % (r->dim[0] * r->dim[1])) / r->dim[0];
        const uint32_t ix = cellLinearIndex % r->dim[0];
        const float3 cellCenter = r->gridOrigin + float3((ix + 0.5f) * r->gridCellSize.x, (iy + 0.5f) * r->gridCellSize.y,
                                                          (iz + 0.5f) * r->gridCellSize.z);
        for (uint32_t slotInCell = 0; slotInCell < kNumLightSlotsPerCell; ++slotInCell) {
            const uint32_t linearThreadIndex = cellLinearIndex * kNumLightSlotsPerCell + slotInCell;
            PCG32RNG rng{ r->slotRngs[linearThreadIndex] };
            float selectedTargetPDensity = 0.0f;
            Reservoir reservoir;
            reservoir.initialize(emptyLightSample());
            for (uint32_t candIdx = 0; candIdx < numCandidates; ++candIdx) {
                float uLight = rng.getFloat0cTo1o();
                bool sampleEnvLight = false;
                float probToSampleCurLightType = 1.0f;
                if (useEnvLight(s, p)) { // build_cell_reservoirs.cu:120-139
                    if (s->instIntegral > 0.0f) {
                        const float prob = std::fmin(std::fmax(kProbToSampleEnvLight * numCandidates - candIdx, 0.0f), 1.0f);
                        if (uLight < prob) {
                            probToSampleCurLightType = kProbToSampleEnvLight;
                            uLight = uLight / prob;
                            sampleEnvLight = true;
                        }
                        else {
                            probToSampleCurLightType = 1.0f - kProbToSampleEnvLight;
                            uLight = (uLight - prob) / (1 - prob);
                        }
                    }
                    else {
                        sampleEnvLight = true;
                    }
                }
                LightSample lightSample = emptyLightSample();
                float areaPDensity = 0.0f;
                const float uPos0 = rng.getFloat0cTo1o();
                const float uPos1 = rng.getFloat0cTo1o();
                const float3 cont = sampleIntensity(s, p, cellCenter, halfCellSize, minSquaredDistance, uLight, sampleEnvLight, uPos0, uPos1,
                                                    &lightSample, &areaPDensity);
                areaPDensity *= probToSampleCurLightType;
                const float targetPDensity = convertToWeight(cont);
                const float weight = targetPDensity / areaPDensity;
                if (reservoir.update(lightSample, weight, rng.getFloat0cTo1o()))
                    selectedTargetPDensity = targetPDensity;
            }
            float recPDFEstimate = reservoir.sumWeights / (selectedTargetPDensity * reservoir.streamLength);
            if (!std::isfinite(recPDFEstimate)) {
                recPDFEstimate = 0.0f;
                selectedTargetPDensity = 0.0f;
            }
            if (useTemporalReuse) {
                const uint32_t selfStreamLength = reservoir.streamLength;
                if (recPDFEstimate == 0.0f)
                    reservoir.initialize(emptyLightSample());
                uint32_t combinedStreamLength = selfStreamLength;
                const uint32_t maxNumPrevSamples = 20 * selfStreamLength;
                const RegirSlot &prevSlot = r->slots[prevBufferIndex][linearThreadIndex];
                const Reservoir prevReservoir = loadSlot(prevSlot);
                const float prevTargetDensity = prevSlot.targetDensity;
                const uint32_t prevStreamLength = std::min(prevReservoir.streamLength, maxNumPrevSamples);
                const float lengthCorrection = static_cast<float>(prevStreamLength) / prevReservoir.streamLength;
                const float weight = lengthCorrection * prevReservoir.sumWeights;
                if (reservoir.update(prevReservoir.sample, weight, rng.getFloat0cTo1o()))
                    selectedTargetPDensity = prevTargetDensity;
                combinedStreamLength += prevStreamLength;
                reservoir.streamLength = combinedStreamLength;
                const float weightForEstimate = 1.0f / reservoir.streamLength;
                recPDFEstimate = weightForEstimate * reservoir.sumWeights / selectedTargetPDensity;
                if (!std::isfinite(recPDFEstimate)) {
                    recPDFEstimate = 0.0f;
                    selectedTargetPDensity = 0.0f;
                }
            }
            r->slotRngs[linearThreadIndex] = rng.state;
            storeSlot(r->slots[bufferIndex][linearThreadIndex], reservoir, recPDFEstimate, selectedTargetPDensity);
        }
    }
}

// build_cell_reservoirs.cu:235-248
extern "C" void orc_regir_update_access(orc_frame* f, const GfxFrameParams* p, uint32_t frameIndex) {
    orc_regir* r = regirState(f, p);
    const uint32_t bufferIndex = p->bufferIndex & 1;
    for (uint32_t cell = 0; cell < r->numCells; ++cell) {
        if (r->perCellNumAccesses[cell] > 0) {
            r->lastAccessFrameIndices[cell] = frameIndex;
            ++r->numActiveCells[bufferIndex];
        }
    }
}

// regir_shared.h:731-742 (float -> uint32 with the GPU's saturating conversion)
static inline uint32_t calcCellLinearIndex(const orc_regir* r, const float3 &positionInWorld) {
    const float3 relPos = positionInWorld - r->gridOrigin;
    const uint32_t ix = std::min(dm_f2uint(relPos.x / r->gridCellSize.x), r->dim[0] - 1);
    const uint32_t iy = std::min(dm_f2uint(relPos.y / r->gridCellSize.y), r->dim[1] - 1);
    const uint32_t iz = std::min(dm_f2uint(relPos.z / r->gridCellSize.z), r->dim[2] - 1);
    return iz * r->dim[0] * r->dim[1] + iy * r->dim[0] + ix;
}

// optix_pathtracing_kernels.cu:18-82
static float3 sampleFromCell(const orc_scene* s, orc_regir* r, const GfxFrameParams* p, const float3 &shadingPoint,
                             const float3 &vOutLocal, const ReferenceFrame &shadingFrame, const BSDF &bsdf, PCG32RNG &rng,
                             LightSample* lightSample, float* recProbDensityEstimate) {
    float3 randomOffset(0.0f);
    if (p->regirEnableCellRandomization) {
        const float o0 = -0.5f + rng.getFloat0cTo1o();
        const float o1 = -0.5f + rng.getFloat0cTo1o();
        const float o2 = -0.5f + rng.getFloat0cTo1o();
        randomOffset = r->gridCellSize * float3(o0, o1, o2);
    }
    const uint32_t cellLinearIndex = calcCellLinearIndex(r, shadingPoint + randomOffset);
    const uint32_t resStartIndex = kNumLightSlotsPerCell * cellLinearIndex;
#pragma omp atomic
    r->perCellNumAccesses[cellLinearIndex] += 1u;

    const uint32_t numResampling = 1u << p->regirLog2NumCandidatesPerCell;
    Reservoir combinedReservoir;
    combinedReservoir.initialize(emptyLightSample());
    uint32_t combinedStreamLength = 0;
    float3 selectedContribution(0.0f);
    float selectedTargetPDensity = 0.0f;
    const uint32_t bufferIndex = p->bufferIndex & 1;
    for (uint32_t i = 0; i < numResampling; ++i) {
        // mapPrimarySampleToDiscrete (common_shared.h:140-150)
        const uint32_t lightSlotIdx = resStartIndex +
            std::min(dm_f2uint(rng.getFloat0cTo1o() * kNumLightSlotsPerCell), kNumLightSlotsPerCell - 1);
        const RegirSlot &slot = r->slots[bufferIndex][lightSlotIdx];
        const Reservoir res = loadSlot(slot);
        const uint32_t streamLength = res.streamLength;
        combinedStreamLength += streamLength;
        if (slot.recPDFEstimate == 0.0f)
            continue;
        const float3 cont = performDirectLighting<false>(s, shadingPoint, vOutLocal, shadingFrame, bsdf, res.sample);
        const float targetPDensity = convertToWeight(cont);
        const float weight = targetPDensity * slot.recPDFEstimate * streamLength;
        if (combinedReservoir.update(res.sample, weight, rng.getFloat0cTo1o())) {
            selectedContribution = cont;
            selectedTargetPDensity = targetPDensity;
        }
    }
    combinedReservoir.streamLength = combinedStreamLength;
    *lightSample = combinedReservoir.sample;
    const float weightForEstimate = 1.0f / combinedReservoir.streamLength;
    *recProbDensityEstimate = weightForEstimate * combinedReservoir.sumWeights / selectedTargetPDensity;
    if (!std::isfinite(*recProbDensityEstimate))
        *recProbDensityEstimate = 0.0f;
    return selectedContribution;
}

// performNextEventEstimation<true> (:84-101)
static float3 regirNextEventEstimation(const orc_scene* s, orc_regir* r, const GfxFrameParams* p, const float3 &shadingPoint,
                                       const float3 &vOutLocal, const ReferenceFrame &shadingFrame, const BSDF &bsdf, PCG32RNG &rng,
                                       PathTraceCounters* counters) {
    float3 ret(0.0f);
    LightSample lightSample;
    float recProbDensityEstimate;
    const float3 unshadowedContribution = sampleFromCell(s, r, p, shadingPoint, vOutLocal, shadingFrame, bsdf, rng, &lightSample,
                                                         &recProbDensityEstimate);
    if (recProbDensityEstimate > 0.0f) {
        ++counters->visibilityRays;
        const float visibility = evaluateVisibility(s, shadingPoint, lightSample) ? 1.0f : 0.0f;
        ret = unshadowedContribution * (visibility * recProbDensityEstimate);
    }
    return ret;
}

static void regirPathTracePixel(orc_frame* f, orc_regir* r, const GfxFrameParams* p, const Camera &camera, uint32_t x, uint32_t y,
                                PathTraceCounters* counters) {
    const orc_scene* s = f->scene;
    const size_t pix = (size_t)y * f->W + x;
    const uint32_t bufIdx = p->bufferIndex & 1;
    const GB0 gb0 = f->gb0[bufIdx][pix];
    const float bcB = decodeBarycentric((uint16_t)(gb0.qbc & 0xFFFFu));
    const float bcC = decodeBarycentric((uint16_t)(gb0.qbc >> 16));

    float3 contribution(0.001f, 0.001f, 0.001f);
    if (gb0.instSlot != 0xFFFFFFFFu) {
        const InstData* inst = &s->instances[gb0.instSlot];
        const MeshData* mesh = &s->meshes[gb0.geomInstSlot];
        SurfacePoint sp;
        computeSurfacePointFromGBuffer(s, *inst, *mesh, gb0.primIndex, bcB, bcC, &sp);

        float3 alpha(1.0f);
        const float initImportance = sRGB_calcLuminance(alpha);
        PCG32RNG rng{ f->rng[pix] };
        float3 positionInWorld = sp.positionInWorld;
        float3 vIn;
        float dirPDensity;
        {
            const GfxMaterialDesc &mat = s->materials[mesh->materialSlot];
            const float3 vOut = normalize(camera.position - positionInWorld);
            const float frontHit = dot(vOut, sp.geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
            positionInWorld = offsetRayOrigin(positionInWorld, frontHit * sp.geometricNormalInWorld);
            const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
            const float3 vOutLocal = shadingFrame.toLocal(vOut);
            contribution = float3(0.0f);
            if (vOutLocal.z > 0 && mat.hasEmittance) {
                const float3 emittance(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
                contribution += alpha * emittance / kPi;
            }
            const BSDF bsdf = setupBsdf(s, mesh->materialSlot, sp.texCoord);
            contribution += alpha * regirNextEventEstimation(s, r, p, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, counters);
            float3 vInLocal;
            const float uDir0 = rng.getFloat0cTo1o();
            const float uDir1 = rng.getFloat0cTo1o();
            alpha *= bsdf.sampleThroughput(vOutLocal, uDir0, uDir1, &vInLocal, &dirPDensity);
            vIn = shadingFrame.fromLocal(vInLocal);
        }

        float prevDirPDensity = dirPDensity;
        uint32_t pathLength = 1;
        float3 rayOrg = positionInWorld;
        float3 rayDir = vIn;
        while (true) {
            const bool isValidSampling = prevDirPDensity > 0.0f && std::isfinite(prevDirPDensity);
            if (!isValidSampling)
                break;
            ++pathLength;
            const bool maxLengthTerminate = pathLength >= (p->maxPathLength ? p->maxPathLength : 5u);
            // useReGIR: length limit and Russian roulette before the trace (:254-263) ...
            if (maxLengthTerminate)
                break;
            {
                const float continueProb = std::fmin(sRGB_calcLuminance(alpha) / initImportance, 1.0f);
                if (rng.getFloat0cTo1o() >= continueProb)
                    break;
                alpha /= continueProb;
            }

            ++counters->closestRays;
            const HitObject hit = traverseCanonical(s->bvh, rayOrg, rayDir, 0.0f, std::numeric_limits<float>::max());
            if (hit.primIndex == UINT32_MAX)
                break;

            inst = &s->instances[s->geomToInst[hit.geomIndex]];
            mesh = &s->meshes[s->geomToMesh[hit.geomIndex]];
            computeSurfacePointAtHit(s, p, *inst, *mesh, hit.primIndex, hit.bcB, hit.bcC, &sp);
            sp.hypAreaPDensity = 0.0f; // never written by computeSurfacePoint<false, ...>: see the header of this file
            const GfxMaterialDesc &mat = s->materials[mesh->materialSlot];

            const float3 vOut = normalize(-rayDir);
            const float frontHit = dot(vOut, sp.geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
            const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
            positionInWorld = offsetRayOrigin(sp.positionInWorld, frontHit * sp.geometricNormalInWorld);
            const float3 vOutLocal = shadingFrame.toLocal(vOut);

            // ... and, because useImplicitLightSampling is on, once more in the closest-hit program (:343-364)
            if (vOutLocal.z > 0 && mat.hasEmittance) {
                const float3 emittance(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
                const float dist2 = sqLength(positionInWorld - rayOrg);
                const float lightPDensity = sp.hypAreaPDensity * dist2 / vOutLocal.z;
                const float bsdfPDensity = prevDirPDensity;
                const float misWeight = pow2(bsdfPDensity) / (pow2(bsdfPDensity) + pow2(lightPDensity));
                contribution += alpha * emittance * (misWeight / kPi);
            }
            const float continueProb = std::fmin(sRGB_calcLuminance(alpha) / initImportance, 1.0f);
            if (rng.getFloat0cTo1o() >= continueProb || maxLengthTerminate)
                break;
            alpha /= continueProb;

            const BSDF bsdf = setupBsdf(s, mesh->materialSlot, sp.texCoord);
            contribution += alpha * regirNextEventEstimation(s, r, p, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, counters);

            float3 vInLocal;
            const float uDir0 = rng.getFloat0cTo1o();
            const float uDir1 = rng.getFloat0cTo1o();
            alpha *= bsdf.sampleThroughput(vOutLocal, uDir0, uDir1, &vInLocal, &dirPDensity);
            rayOrg = positionInWorld;
            rayDir = shadingFrame.fromLocal(vInLocal);
            prevDirPDensity = dirPDensity;
        }
        f->rng[pix] = rng.state;
    }
    else if (useEnvLight(s, p)) { // regir/gpu_kernels/optix_pathtracing_kernels.cu:283-290; the ReGIR ray type's miss program is empty
        contribution = p->envLightPowerCoeff * s->env.fetch(bcB, bcC);
    }

    float3 prevColorResult(0.0f);
    if (p->numAccumFrames > 0)
        prevColorResult = float3(f->beauty[pix].x, f->beauty[pix].y, f->beauty[pix].z);
    const float curWeight = 1.0f / (1 + p->numAccumFrames);
    const float3 colorResult = (1 - curWeight) * prevColorResult + curWeight * contribution;
    f->beauty[pix] = F4{ colorResult.x, colorResult.y, colorResult.z, 1.0f };
}

static uint64_t regirPathTrace(orc_frame* f, const GfxFrameParams* p, int numThreads) {
    orc_regir* r = regirState(f, p);
    const Camera camera = makeCamera(p->camera);
    const uint32_t W = f->W, H = f->H;
    const uint32_t y0 = p->tileOriginY, y1 = p->tileRows ? std::min(H, p->tileOriginY + p->tileRows) : H;
    uint64_t rays = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(numThreads) reduction(+ : rays)
    for (int64_t yy = y0; yy < (int64_t)y1; ++yy) {
        PathTraceCounters counters;
        for (uint32_t x = 0; x < W; ++x)
            regirPathTracePixel(f, r, p, camera, x, (uint32_t)yy, &counters);
        rays += counters.closestRays + counters.visibilityRays;
    }
    return rays;
}

static void* regirBufferPtr(orc_frame* f, int id, uint32_t index, size_t* bytes) {
    orc_regir* r = f->regir;
    void* ptr = nullptr;
    size_t b = 0;
    if (r) {
        switch (id) {
        case GFX_BUF_REGIR_SLOTS: ptr = r->slots[index & 1].data(); b = (size_t)r->numSlots * 64; break;
        case GFX_BUF_REGIR_SLOT_RNG: ptr = r->slotRngs.data(); b = (size_t)r->numSlots * 8; break;
        case GFX_BUF_REGIR_CELL_ACCESSES: ptr = r->perCellNumAccesses.data(); b = (size_t)r->numCells * 4; break;
        case GFX_BUF_REGIR_LAST_ACCESS: ptr = r->lastAccessFrameIndices.data(); b = (size_t)r->numCells * 4; break;
        case GFX_BUF_REGIR_NUM_ACTIVE_CELLS: ptr = r->numActiveCells; b = 8; break;
        default: break;
        }
    }
    if (bytes) *bytes = b;
    return ptr;
}
