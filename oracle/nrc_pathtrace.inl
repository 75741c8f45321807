// oracle/nrc_pathtrace.inl — TEST INFRASTRUCTURE (CPU oracle), included by render.cpp after pathtrace.inl.
//
// CPU restatement of the Neural Radiance Caching path tracer and its bookkeeping kernels:
//   convertToPolar / createRadianceQuery       neural_radiance_caching/gpu_kernels/optix_pathtracing_kernels.cu:12-34
//   pathTrace_raygen_generic<true>             :95-361
//   pathTrace_closestHit_generic<true>         :363-623        (miss :625-670 is a no-op without an env light)
//   preprocessNRC                              neural_radiance_caching/gpu_kernels/nrc_setup_kernels.cu:6-49
//   accumulateInferredRadianceValues           :51-92
//   propagateRadianceValues                    :94-138
//   shuffleTrainingData                        :140-216
//   LinearCongruentialGenerator + seeding      neural_radiance_caching_shared.h:164-181, neural_radiance_caching_main.cpp:1187-1193
// constants: pathTerminationFactor 0.01, numTrainingDataPerFrame 65536, trainBufferSize 131072,
// useReflectanceFactorization = true (neural_radiance_caching_shared.h:7-10).
//
// Allocation order of training vertices.  The reference allocates records with a global atomicAdd
// (:216, :571), so the index a vertex receives depends on GPU scheduling: every order is a valid execution.
// Oracle and product both use the canonical order "path vertex number first, tile index second": the
// paths are advanced in lock step, one vertex per round, and after each round the vertices staged by the
// training paths are numbered in tile order.  That makes frames reproducible bit for bit (and identical on
// every GPU of a multi-GPU run, which must keep the network weights in sync).

static constexpr float kPathTerminationFactor = 0.01f;
static constexpr uint32_t kNumTrainingDataPerFrame = 1u << 16;
static constexpr uint32_t kTrainBufferSize = 2 * kNumTrainingDataPerFrame;
static constexpr uint32_t kInvalidVertexDataIndex = 0x007FFFFFu;
static constexpr uint32_t kMaxNrcRounds = 62; // PathTraceReadWritePayload::pathLength is a 6-bit field

// indices into the NRC state block (GFX_BUF_NRC_STATE)
enum {
    NRC_NUM_TRAINING_DATA = 0,   // [2]
    NRC_TILE_SIZE = 2,           // [2][2]
    NRC_OFFSET_UNBIASED_TILE = 6,
    NRC_OFFSET_TRAINING_PATH = 7,
    NRC_TARGET_MIN = 8,          // [2][3] ordered ints: buffer 0 at 8, buffer 1 at 14
    NRC_TARGET_MAX = 11,         //                      buffer 0 at 11, buffer 1 at 17
    NRC_TARGET_AVG = 20,         // [2][3] floats
    NRC_NUM_INFERENCE_QUERIES = 26,
    NRC_STATE_WORDS = 32
};

struct NrcPathState { // PathTraceReadWritePayload<true> + the ray (neural_radiance_caching_shared.h:214-236)
    bool active = false;
    PCG32RNG rng;
    float3 alpha, contribution;
    float prevDirPDensity;
    uint32_t linearTileIndex;
    float primaryPathSpread, curSqrtPathSpread;
    float3 prevLocalThroughput;
    uint32_t prevTrainDataIndex;
    bool renderingPathEndsWithCache, isTrainingPath, isUnbiasedTrainingTile, trainingSuffixEndsWithCache;
    uint32_t pathLength;
    float3 rayOrg, rayDir;
};

struct NrcStagedVertex { // a training vertex waiting for its index
    bool want = false;
    bool fromRayGen = false;
    float query[14];
    float3 localThroughput, directContNEE;
    uint32_t pathLength;
};

struct orc_nrc_frame {
    uint32_t maxNumTrainingSuffixes;
    std::vector<float> inferenceQuery;      // 14 x (W*H + maxNumTrainingSuffixes)
    std::vector<uint32_t> terminalInfo;     // 4 x W*H: alpha rgb (float bits), hasQuery | pathLength<<1 | isTrainingPixel<<9 | isUnbiasedTile<<10
    std::vector<float> inferredRadiance;    // 3 x (W*H + maxNumTrainingSuffixes)
    std::vector<float> frameContribution;   // 3 x W*H
    std::vector<float> trainQuery[2];       // 14 x trainBufferSize
    std::vector<float> trainTarget[2];      // 3 x trainBufferSize
    std::vector<uint32_t> trainVertexInfo;  // 4 x trainBufferSize: localThroughput rgb, prev | pathLength<<23
    std::vector<uint32_t> suffixTerminal;   // maxNumTrainingSuffixes: prev | hasQuery<<23 | pathLength<<24
    std::vector<uint32_t> shufflers;        // numTrainingDataPerFrame LCG states
    uint32_t state[NRC_STATE_WORDS];
    std::vector<NrcPathState> paths;
    std::vector<NrcStagedVertex> staged;
    // strip sharding over ranks, the restatement of gfx_nrc_shard for the gloo tests of the host logic (tests/test_multigpu_cpu.py):
    // every rank traces its rows; per commit round the ranks exchange (vertex count, any path still active), which makes the
    // record numbering - tile order = rank order, then local tile order - that of the unsharded frame; after the propagation
    // the records are merged by an unsigned integer sum over zero-initialised arrays.  The collectives are the caller's.
    int shardRank = 0, shardWorld = 1;
    void (*shardExchange)(void* user, const uint32_t* mine, uint32_t numWords, uint32_t* all) = nullptr;
    void (*shardSum)(void* user, uint32_t* words, uint64_t numWords) = nullptr;
    void* shardUser = nullptr;
    bool sharded() const { return shardWorld > 1 && shardExchange && shardSum; }
};

static orc_nrc_frame* nrcFrame(orc_frame* f) {
    if (f->nrc)
        return f->nrc;
    orc_nrc_frame* n = new orc_nrc_frame();
    const size_t numPixels = (size_t)f->W * f->H;
    // W*H/16 in the reference (neural_radiance_caching_main.cpp:1151); rounded up here so that 4x4 tiles of an image
    // whose size is not a multiple of 4 still have a slot (the reference would index out of bounds)
    n->maxNumTrainingSuffixes = ((f->W + 3) / 4) * ((f->H + 3) / 4);
    const size_t queryCapacity = (numPixels + n->maxNumTrainingSuffixes + 127) / 128 * 128;
    n->inferenceQuery.assign(14 * queryCapacity, 0.0f);
    n->terminalInfo.assign(4 * numPixels, 0u);
    n->inferredRadiance.assign(3 * queryCapacity, 0.0f);
    n->frameContribution.assign(3 * numPixels, 0.0f);
    for (int i = 0; i < 2; ++i) {
        n->trainQuery[i].assign(14 * (size_t)kTrainBufferSize, 0.0f);
        n->trainTarget[i].assign(3 * (size_t)kTrainBufferSize, 0.0f);
    }
    n->trainVertexInfo.assign(4 * (size_t)kTrainBufferSize, 0u);
    n->suffixTerminal.assign(n->maxNumTrainingSuffixes, kInvalidVertexDataIndex);
    // neural_radiance_caching_main.cpp:1187-1193
    n->shufflers.resize(kNumTrainingDataPerFrame);
    uint32_t lcg = 471313181u;
    for (uint32_t i = 0; i < kNumTrainingDataPerFrame; ++i) {
        lcg = (lcg * 1103515245u + 12345u) % (1u << 31);
        n->shufflers[i] = lcg;
    }
    std::memset(n->state, 0, sizeof(n->state));
    n->state[NRC_TILE_SIZE + 0] = n->state[NRC_TILE_SIZE + 1] = 8; // tileSize[i].initialize(.., uint2(8, 8)) :1160
    n->state[NRC_TILE_SIZE + 2] = n->state[NRC_TILE_SIZE + 3] = 8;
    n->paths.resize(numPixels);
    n->staged.resize(n->maxNumTrainingSuffixes);
    f->nrc = n;
    return n;
}

static void nrcFrameDestroy(orc_nrc_frame* n) { delete n; }

static void* nrcBufferPtr(orc_frame* f, int id, uint32_t index, size_t* bytes) {
    orc_nrc_frame* n = nrcFrame(f);
    void* p = nullptr;
    size_t b = 0;
    switch (id) {
    case GFX_BUF_NRC_INFERENCE_QUERY: p = n->inferenceQuery.data(); b = n->inferenceQuery.size() * 4; break;
    case GFX_BUF_NRC_TERMINAL_INFO: p = n->terminalInfo.data(); b = n->terminalInfo.size() * 4; break;
    case GFX_BUF_NRC_INFERRED_RADIANCE: p = n->inferredRadiance.data(); b = n->inferredRadiance.size() * 4; break;
    case GFX_BUF_NRC_FRAME_CONTRIBUTION: p = n->frameContribution.data(); b = n->frameContribution.size() * 4; break;
    case GFX_BUF_NRC_TRAIN_QUERY: p = n->trainQuery[index & 1].data(); b = n->trainQuery[index & 1].size() * 4; break;
    case GFX_BUF_NRC_TRAIN_TARGET: p = n->trainTarget[index & 1].data(); b = n->trainTarget[index & 1].size() * 4; break;
    case GFX_BUF_NRC_TRAIN_VERTEX_INFO: p = n->trainVertexInfo.data(); b = n->trainVertexInfo.size() * 4; break;
    case GFX_BUF_NRC_TRAIN_SUFFIX_TERMINAL: p = n->suffixTerminal.data(); b = n->suffixTerminal.size() * 4; break;
    case GFX_BUF_NRC_STATE: p = n->state; b = sizeof(n->state); break;
    default: break;
    }
    if (bytes) *bytes = b;
    return p;
}

static inline int32_t floatToOrderedInt(float v) { // basic_types.h:411-418
    const int32_t i = (int32_t)f2u(v);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}

// optix_pathtracing_kernels.cu:12-16
static inline void convertToPolar(const float3 &dir, float* phi, float* theta) {
    const float z = std::fmin(std::fmax(dir.z, -1.0f), 1.0f);
    *theta = dm_acos(z);
    *phi = dm_atan2(dir.y, dir.x);
}

// optix_pathtracing_kernels.cu:18-34; AABB::normalize = safeDivide(p - minP, maxP - minP) (basic_types.h:3425-3427)
static void createRadianceQuery(const GfxFrameParams* p, const float3 &positionInWorld, const float3 &normalInWorld,
                                const float3 &scatteredDirInWorld, const BSDF &bsdf, float* q) {
    const float3 minP(p->sceneAabbMin[0], p->sceneAabbMin[1], p->sceneAabbMin[2]);
    const float3 maxP(p->sceneAabbMax[0], p->sceneAabbMax[1], p->sceneAabbMax[2]);
    const float3 a = positionInWorld - minP, d = maxP - minP;
    q[0] = d.x != 0 ? a.x / d.x : 0.0f;
    q[1] = d.y != 0 ? a.y / d.y : 0.0f;
    q[2] = d.z != 0 ? a.z / d.z : 0.0f;
    convertToPolar(normalInWorld, &q[3], &q[4]);
    convertToPolar(scatteredDirInWorld, &q[5], &q[6]);
    // BSDF::getSurfaceParameters (common_device.cuh:342-347, 525-531)
    q[7] = 1 - dm_exp(-bsdf.roughness);
    q[8] = bsdf.diffuseColor.x; q[9] = bsdf.diffuseColor.y; q[10] = bsdf.diffuseColor.z;
    q[11] = bsdf.specularF0Color.x; q[12] = bsdf.specularF0Color.y; q[13] = bsdf.specularF0Color.z;
}

static inline void writeTerminalInfo(orc_nrc_frame* n, size_t pix, const float3 &alpha, uint32_t pathLength, bool hasQuery,
                                     bool isTrainingPixel, bool isUnbiasedTile) { // TerminalInfo, neural_radiance_caching_shared.h:139-146
    uint32_t* t = &n->terminalInfo[4 * pix];
    t[0] = f2u(alpha.x); t[1] = f2u(alpha.y); t[2] = f2u(alpha.z);
    t[3] = (hasQuery ? 1u : 0u) | ((pathLength & 0xFFu) << 1) | ((isTrainingPixel ? 1u : 0u) << 9) | ((isUnbiasedTile ? 1u : 0u) << 10);
}
static inline uint32_t packSuffixTerminal(uint32_t prev, bool hasQuery, uint32_t pathLength) { // :158-163
    return (prev & 0x7FFFFFu) | ((hasQuery ? 1u : 0u) << 23) | ((pathLength & 0xFFu) << 24);
}

// nrc_setup_kernels.cu:6-49
extern "C" void orc_nrc_set_shard(orc_frame* f, int rank, int world,
                                  void (*exchange)(void*, const uint32_t*, uint32_t, uint32_t*),
                                  void (*sum)(void*, uint32_t*, uint64_t), void* user) {
    orc_nrc_frame* n = nrcFrame(f);
    n->shardRank = rank;
    n->shardWorld = world;
    n->shardExchange = exchange;
    n->shardSum = sum;
    n->shardUser = user;
}

extern "C" void orc_nrc_preprocess(orc_frame* f, const GfxFrameParams* p, uint32_t offsetToSelectUnbiasedTile,
                                   uint32_t offsetToSelectTrainingPath, int isNewSequence) {
    orc_nrc_frame* n = nrcFrame(f);
    if (n->sharded()) { // the ranks' records are merged by an integer sum after the propagation: start from zero
        std::fill(n->trainQuery[0].begin(), n->trainQuery[0].end(), 0.0f);
        std::fill(n->trainTarget[0].begin(), n->trainTarget[0].end(), 0.0f);
    }
    const uint32_t bufIdx = p->bufferIndex & 1, prevBufIdx = (bufIdx + 1) % 2;
    uint32_t newTileSize[2];
    if (isNewSequence) {
        newTileSize[0] = newTileSize[1] = 8;
    }
    else {
        const uint32_t prevNumTrainingData = n->state[NRC_NUM_TRAINING_DATA + prevBufIdx];
        const float r = std::sqrt(static_cast<float>(prevNumTrainingData) / kNumTrainingDataPerFrame);
        for (int c = 0; c < 2; ++c) {
            const uint32_t cur = n->state[NRC_TILE_SIZE + 2 * prevBufIdx + c];
            newTileSize[c] = std::min(std::max(dm_f2uint(cur * r), 4u), 128u);
        }
    }
    n->state[NRC_TILE_SIZE + 2 * bufIdx + 0] = newTileSize[0];
    n->state[NRC_TILE_SIZE + 2 * bufIdx + 1] = newTileSize[1];
    n->state[NRC_NUM_TRAINING_DATA + bufIdx] = 0;
    n->state[NRC_OFFSET_UNBIASED_TILE] = offsetToSelectUnbiasedTile;
    n->state[NRC_OFFSET_TRAINING_PATH] = offsetToSelectTrainingPath;
    const float inf = std::numeric_limits<float>::infinity();
    for (int c = 0; c < 3; ++c) {
        n->state[NRC_TARGET_MIN + 6 * bufIdx + c] = (uint32_t)floatToOrderedInt(inf);
        n->state[NRC_TARGET_MAX + 6 * bufIdx + c] = (uint32_t)floatToOrderedInt(-inf);
        n->state[NRC_TARGET_AVG + 3 * bufIdx + c] = f2u(0.0f);
    }
    // what the host computes after reading tileSize back (neural_radiance_caching_main.cpp:2301-2304)
    const uint32_t numTilesX = (f->W + newTileSize[0] - 1) / newTileSize[0], numTilesY = (f->H + newTileSize[1] - 1) / newTileSize[1];
    n->state[NRC_NUM_INFERENCE_QUERIES] = (f->W * f->H + numTilesX * numTilesY + 127) / 128 * 128;
    for (uint32_t i = 0; i < n->maxNumTrainingSuffixes; ++i)
        n->suffixTerminal[i] = packSuffixTerminal(kInvalidVertexDataIndex, false, 0);
}

// ray generation up to the path extension loop (:95-283)
static void nrcRayGen(orc_frame* f, orc_nrc_frame* n, const GfxFrameParams* p, const Camera &camera, uint32_t x, uint32_t y,
                      PathTraceCounters* counters) {
    const orc_scene* s = f->scene;
    const size_t pix = (size_t)y * f->W + x;
    const uint32_t bufIdx = p->bufferIndex & 1;
    const GB0 gb0 = f->gb0[bufIdx][pix];
    const float bcB = decodeBarycentric((uint16_t)(gb0.qbc & 0xFFFFu));
    const float bcC = decodeBarycentric((uint16_t)(gb0.qbc >> 16));
    NrcPathState &st = n->paths[pix];
    st = NrcPathState();

    const uint32_t tileSizeX = n->state[NRC_TILE_SIZE + 2 * bufIdx], tileSizeY = n->state[NRC_TILE_SIZE + 2 * bufIdx + 1];
    const uint32_t numPixelsInTile = tileSizeX * tileSizeY;
    const uint32_t localLinearIndex = (y % tileSizeY) * tileSizeX + (x % tileSizeX);
    st.isTrainingPath = (localLinearIndex + n->state[NRC_OFFSET_TRAINING_PATH]) % numPixelsInTile == 0;
    const uint32_t numTilesX = (f->W + tileSizeX - 1) / tileSizeX;
    const uint32_t tileX = x / tileSizeX, tileY = y / tileSizeY;
    st.linearTileIndex = tileY * numTilesX + tileX;
    const uint32_t localLinearTileIndex = (tileY % 4) * 4 + (tileX % 4);
    st.isUnbiasedTrainingTile = (localLinearTileIndex + n->state[NRC_OFFSET_UNBIASED_TILE]) % 16 == 0;

    st.contribution = float3(0.001f, 0.001f, 0.001f);
    st.renderingPathEndsWithCache = false;
    st.trainingSuffixEndsWithCache = false;
    st.pathLength = 1;
    if (gb0.instSlot == 0xFFFFFFFFu) {
        if (useEnvLight(s, p)) // :323-331: the environment seen directly, (u, v) left in the barycentrics by the miss program
            st.contribution = p->envLightPowerCoeff * s->env.fetch(bcB, bcC);
        return;
    }

    const InstData &inst = s->instances[gb0.instSlot];
    const MeshData &mesh = s->meshes[gb0.geomInstSlot];
    SurfacePoint sp;
    computeSurfacePointFromGBuffer(s, inst, mesh, gb0.primIndex, bcB, bcC, &sp);

    float3 alpha(1.0f);
    PCG32RNG rng{ f->rng[pix] };
    const GfxMaterialDesc &mat = s->materials[mesh.materialSlot];

    float3 vOut = camera.position - sp.positionInWorld;
    const float primaryDist2 = sqLength(vOut);
    vOut /= std::sqrt(primaryDist2);
    const float primaryDotVN = dot(vOut, sp.geometricNormalInWorld);
    const float frontHit = primaryDotVN >= 0.0f ? 1.0f : -1.0f;
    const float3 positionInWorld = offsetRayOrigin(sp.positionInWorld, frontHit * sp.geometricNormalInWorld);
    st.primaryPathSpread = primaryDist2 / (4 * kPi * std::fabs(primaryDotVN));

    const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
    const float3 vOutLocal = shadingFrame.toLocal(vOut);

    float3 contribution(0.0f);
    if (vOutLocal.z > 0 && mat.hasEmittance) {
        const float3 emittance(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
        contribution += alpha * emittance / kPi;
    }
    const BSDF bsdf = setupBsdf(s, mesh.materialSlot, sp.texCoord);
    const float3 directContNEE = performNextEventEstimation(s, p, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, counters);
    contribution += alpha * directContNEE;

    float3 vInLocal;
    float dirPDensity;
    const float uDir0 = rng.getFloat0cTo1o();
    const float uDir1 = rng.getFloat0cTo1o();
    const float3 localThroughput = bsdf.sampleThroughput(vOutLocal, uDir0, uDir1, &vInLocal, &dirPDensity);
    alpha *= localThroughput;

    st.prevTrainDataIndex = kInvalidVertexDataIndex;
    if (st.isTrainingPath) {
        NrcStagedVertex &sv = n->staged[st.linearTileIndex];
        sv.want = true;
        sv.fromRayGen = true;
        createRadianceQuery(p, positionInWorld, shadingFrame.normal, vOut, bsdf, sv.query);
        sv.localThroughput = localThroughput;
        sv.directContNEE = directContNEE;
        sv.pathLength = st.pathLength;
    }

    st.active = true;
    st.rng = rng;
    st.alpha = alpha;
    st.contribution = contribution;
    st.prevDirPDensity = dirPDensity;
    st.curSqrtPathSpread = 0.0f;
    st.prevLocalThroughput = localThroughput;
    st.rayOrg = positionInWorld;
    st.rayDir = shadingFrame.fromLocal(vInLocal);
}

// one iteration of the path extension loop (:285-311) with the closest-hit program (:363-623) inlined
static void nrcExtend(orc_frame* f, orc_nrc_frame* n, const GfxFrameParams* p, size_t pix, PathTraceCounters* counters) {
    const orc_scene* s = f->scene;
    NrcPathState &st = n->paths[pix];
    const bool isValidSampling = st.prevDirPDensity > 0.0f && std::isfinite(st.prevDirPDensity);
    if (!isValidSampling) {
        st.active = false;
        return;
    }
    ++st.pathLength;
    const bool maxLengthTerminate = (st.pathLength >= p->maxPathLength && p->maxPathLength > 0) || st.pathLength >= kMaxNrcRounds;
    st.active = false; // rwPayload.terminate = true

    ++counters->closestRays;
    const HitObject hit = traverseCanonical(s->bvh, st.rayOrg, st.rayDir, 0.0f, std::numeric_limits<float>::max());
    if (hit.primIndex == UINT32_MAX) { // pathTrace_miss_generic (:625-672)
        if (useEnvLight(s, p)) {
            const float3 directContImplicit = evaluateEnvLightOnMiss(s, p, st.rayDir, st.prevDirPDensity, true);
            st.contribution += st.alpha * directContImplicit;
            if (st.isTrainingPath && st.prevTrainDataIndex != kInvalidVertexDataIndex) {
                float* tgt = &n->trainTarget[0][3 * (size_t)st.prevTrainDataIndex];
                const float3 add = st.prevLocalThroughput * directContImplicit;
                tgt[0] += add.x; tgt[1] += add.y; tgt[2] += add.z;
            }
        }
        return;
    }

    const InstData &inst = s->instances[s->geomToInst[hit.geomIndex]];
    const MeshData &mesh = s->meshes[s->geomToMesh[hit.geomIndex]];
    SurfacePoint sp;
    computeSurfacePointAtHit(s, p, inst, mesh, hit.primIndex, hit.bcB, hit.bcC, &sp);
    const GfxMaterialDesc &mat = s->materials[mesh.materialSlot];

    const float3 vOut = normalize(-st.rayDir);
    const float frontHit = dot(vOut, sp.geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
    const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
    const float3 positionInWorld = offsetRayOrigin(sp.positionInWorld, frontHit * sp.geometricNormalInWorld);
    const float3 vOutLocal = shadingFrame.toLocal(vOut);

    const float dist2 = sqLength(positionInWorld - st.rayOrg);
    st.curSqrtPathSpread += std::sqrt(dist2 / (st.prevDirPDensity * std::fabs(vOutLocal.z)));

    // implicit light sampling (:417-445)
    if (vOutLocal.z > 0 && mat.hasEmittance) {
        const float3 emittance(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
        const float lightPDensity = sp.hypAreaPDensity * dist2 / vOutLocal.z;
        const float bsdfPDensity = st.prevDirPDensity;
        const float misWeight = pow2(bsdfPDensity) / (pow2(bsdfPDensity) + pow2(lightPDensity));
        const float3 directContImplicit = emittance * (misWeight / kPi);
        st.contribution += st.alpha * directContImplicit;
        if (st.isTrainingPath && st.prevTrainDataIndex != kInvalidVertexDataIndex) {
            float* tgt = &n->trainTarget[0][3 * (size_t)st.prevTrainDataIndex];
            const float3 add = st.prevLocalThroughput * directContImplicit;
            tgt[0] += add.x; tgt[1] += add.y; tgt[2] += add.z;
        }
    }

    // Russian roulette (:447-469)
    bool performRR = true;
    bool terminatedByRR = false;
    float recContinueProb = 1.0f;
    if (st.isTrainingPath)
        performRR = st.pathLength > 2;
    if (performRR) {
        const float continueProb = std::fmin(sRGB_calcLuminance(st.alpha) / sRGB_calcLuminance(float3(1.0f)), 1.0f);
        if (st.rng.getFloat0cTo1o() >= continueProb || maxLengthTerminate) {
            if (st.renderingPathEndsWithCache && st.isTrainingPath && st.isUnbiasedTrainingTile)
                return;
            terminatedByRR = true;
        }
        recContinueProb = 1.0f / continueProb;
    }

    const BSDF bsdf = setupBsdf(s, mesh.materialSlot, sp.texCoord);

    // path termination by the spread heuristic (:474-531)
    bool endsWithCache = pow2(st.curSqrtPathSpread) > kPathTerminationFactor * st.primaryPathSpread;
    if (st.renderingPathEndsWithCache && st.isTrainingPath && st.isUnbiasedTrainingTile)
        endsWithCache = false;
    if (endsWithCache) {
        float query[14];
        createRadianceQuery(p, positionInWorld, shadingFrame.normal, vOut, bsdf, query);
        if (!st.renderingPathEndsWithCache) {
            std::memcpy(&n->inferenceQuery[14 * pix], query, sizeof(query));
            writeTerminalInfo(n, pix, st.alpha, st.pathLength, true, st.isTrainingPath, st.isUnbiasedTrainingTile);
            st.renderingPathEndsWithCache = true;
            if (st.isTrainingPath)
                st.curSqrtPathSpread = 0;
            else
                return;
        }
        else {
            if (!st.trainingSuffixEndsWithCache) {
                const size_t offset = (size_t)f->W * f->H;
                std::memcpy(&n->inferenceQuery[14 * (offset + st.linearTileIndex)], query, sizeof(query));
                n->suffixTerminal[st.linearTileIndex] = packSuffixTerminal(st.prevTrainDataIndex, true, st.pathLength);
                st.trainingSuffixEndsWithCache = true;
            }
            return;
        }
    }

    if (terminatedByRR)
        return;
    st.alpha *= recContinueProb;
    if (st.isTrainingPath && st.prevTrainDataIndex != kInvalidVertexDataIndex) {
        uint32_t* vi = &n->trainVertexInfo[4 * (size_t)st.prevTrainDataIndex];
        for (int c = 0; c < 3; ++c)
            vi[c] = f2u(u2f(vi[c]) * recContinueProb);
    }

    const float3 directContNEE = performNextEventEstimation(s, p, positionInWorld, vOutLocal, shadingFrame, bsdf, st.rng, counters);
    st.contribution += st.alpha * directContNEE;

    float3 vInLocal;
    float dirPDensity;
    const float uDir0 = st.rng.getFloat0cTo1o();
    const float uDir1 = st.rng.getFloat0cTo1o();
    const float3 localThroughput = bsdf.sampleThroughput(vOutLocal, uDir0, uDir1, &vInLocal, &dirPDensity);
    st.alpha *= localThroughput;
    st.rayOrg = positionInWorld;
    st.rayDir = shadingFrame.fromLocal(vInLocal);
    st.prevDirPDensity = dirPDensity;
    st.prevLocalThroughput = localThroughput;
    st.active = true; // rwPayload.terminate = false

    if (st.isTrainingPath && !st.trainingSuffixEndsWithCache) {
        NrcStagedVertex &sv = n->staged[st.linearTileIndex];
        sv.want = true;
        sv.fromRayGen = false;
        createRadianceQuery(p, positionInWorld, shadingFrame.normal, vOut, bsdf, sv.query);
        sv.localThroughput = localThroughput;
        sv.directContNEE = directContNEE;
        sv.pathLength = st.pathLength;
    }
}

// numbers the training vertices staged in this round in tile order (:213-248, :568-617); returns whether any path of any rank
// is still active (`localActive` is this rank's answer)
static bool nrcCommitStagedVertices(orc_frame* f, orc_nrc_frame* n, const GfxFrameParams* p, const std::vector<uint32_t> &tileToPixel,
                                    bool localActive = false) {
    const uint32_t bufIdx = p->bufferIndex & 1;
    const size_t offset = (size_t)f->W * f->H;
    uint32_t next = n->state[NRC_NUM_TRAINING_DATA + bufIdx];
    uint32_t counterAfter = 0;
    bool anyActive = localActive;
    if (n->sharded()) {
        uint32_t mine[2] = { 0u, localActive ? 1u : 0u };
        for (uint32_t tile = 0; tile < (uint32_t)tileToPixel.size(); ++tile)
            mine[0] += n->staged[tile].want ? 1u : 0u;
        std::vector<uint32_t> all(2 * (size_t)n->shardWorld);
        n->shardExchange(n->shardUser, mine, 2, all.data());
        uint32_t total = 0;
        anyActive = false;
        for (int r = 0; r < n->shardWorld; ++r) {
            if (r < n->shardRank)
                next += all[2 * r];
            total += all[2 * r];
            anyActive = anyActive || all[2 * r + 1] != 0;
        }
        counterAfter = n->state[NRC_NUM_TRAINING_DATA + bufIdx] + total;
    }
    for (uint32_t tile = 0; tile < (uint32_t)tileToPixel.size(); ++tile) {
        NrcStagedVertex &sv = n->staged[tile];
        if (!sv.want)
            continue;
        sv.want = false;
        NrcPathState &st = n->paths[tileToPixel[tile]];
        const uint32_t trainDataIndex = next++;
        if (trainDataIndex < kTrainBufferSize) {
            std::memcpy(&n->trainQuery[0][14 * (size_t)trainDataIndex], sv.query, sizeof(sv.query));
            uint32_t* vi = &n->trainVertexInfo[4 * (size_t)trainDataIndex];
            vi[0] = f2u(sv.localThroughput.x); vi[1] = f2u(sv.localThroughput.y); vi[2] = f2u(sv.localThroughput.z);
            vi[3] = (st.prevTrainDataIndex & 0x7FFFFFu) | ((sv.pathLength & 0xFFu) << 23);
            float* tgt = &n->trainTarget[0][3 * (size_t)trainDataIndex];
            tgt[0] = sv.directContNEE.x; tgt[1] = sv.directContNEE.y; tgt[2] = sv.directContNEE.z;
            st.prevTrainDataIndex = trainDataIndex;
        }
        else if (sv.fromRayGen) {
            st.prevTrainDataIndex = kInvalidVertexDataIndex; // :244-246
        }
        else { // the buffer is full: end the training suffix here with a query (:605-616)
            std::memcpy(&n->inferenceQuery[14 * (offset + tile)], sv.query, sizeof(sv.query));
            n->suffixTerminal[tile] = packSuffixTerminal(st.prevTrainDataIndex, true, sv.pathLength);
            st.trainingSuffixEndsWithCache = true;
        }
    }
    n->state[NRC_NUM_TRAINING_DATA + bufIdx] = n->sharded() ? counterAfter : next;
    return anyActive;
}

static uint64_t nrcPathTrace(orc_frame* f, const GfxFrameParams* p, int numThreads) {
    orc_nrc_frame* n = nrcFrame(f);
    const Camera camera = makeCamera(p->camera);
    const uint32_t W = f->W, H = f->H;
    const uint32_t bufIdx = p->bufferIndex & 1;
    const uint32_t tileSizeX = n->state[NRC_TILE_SIZE + 2 * bufIdx], tileSizeY = n->state[NRC_TILE_SIZE + 2 * bufIdx + 1];
    const uint32_t numTilesX = (W + tileSizeX - 1) / tileSizeX, numTilesY = (H + tileSizeY - 1) / tileSizeY;
    std::vector<uint32_t> tileToPixel((size_t)numTilesX * numTilesY, 0u);
    for (auto &sv : n->staged)
        sv.want = false;

    // a strip of rows (tileOriginY / tileRows) is one rank's share of a sharded frame
    const uint32_t y0 = p->tileOriginY, y1 = p->tileRows ? std::min(H, p->tileOriginY + p->tileRows) : H;
    for (size_t pix = 0; pix < (size_t)W * H; ++pix)
        if (pix < (size_t)y0 * W || pix >= (size_t)y1 * W)
            n->paths[pix].active = false;
    uint64_t rays = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(numThreads) reduction(+ : rays)
    for (int64_t y = y0; y < (int64_t)y1; ++y) {
        PathTraceCounters counters;
        for (uint32_t x = 0; x < W; ++x) {
            nrcRayGen(f, n, p, camera, x, (uint32_t)y, &counters);
            const NrcPathState &st = n->paths[(size_t)y * W + x];
            if (st.isTrainingPath)
                tileToPixel[st.linearTileIndex] = (uint32_t)(y * W + x);
        }
        rays += counters.closestRays + counters.visibilityRays;
    }
    nrcCommitStagedVertices(f, n, p, tileToPixel);

    for (uint32_t round = 0; round < kMaxNrcRounds; ++round) {
        uint64_t numActive = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(numThreads) reduction(+ : rays, numActive)
        for (int64_t y = y0; y < (int64_t)y1; ++y) {
            PathTraceCounters counters;
            for (uint32_t x = 0; x < W; ++x) {
                const size_t pix = (size_t)y * W + x;
                if (!n->paths[pix].active)
                    continue;
                nrcExtend(f, n, p, pix, &counters);
                numActive += n->paths[pix].active ? 1 : 0;
            }
            rays += counters.closestRays + counters.visibilityRays;
        }
        if (!nrcCommitStagedVertices(f, n, p, tileToPixel, numActive != 0))
            break;
    }

    // ray-gen epilogue (:312-360)
    for (size_t pix = (size_t)y0 * W; pix < (size_t)y1 * W; ++pix) {
        const NrcPathState &st = n->paths[pix];
        if (f->gb0[bufIdx][pix].instSlot != 0xFFFFFFFFu) {
            f->rng[pix] = st.rng.state;
            if (st.isTrainingPath && !st.trainingSuffixEndsWithCache)
                n->suffixTerminal[st.linearTileIndex] = packSuffixTerminal(st.prevTrainDataIndex, false, st.pathLength);
        }
        if (!st.renderingPathEndsWithCache)
            writeTerminalInfo(n, pix, float3(0.0f), st.pathLength, false, st.isTrainingPath, st.isUnbiasedTrainingTile);
        n->frameContribution[3 * pix + 0] = st.contribution.x;
        n->frameContribution[3 * pix + 1] = st.contribution.y;
        n->frameContribution[3 * pix + 2] = st.contribution.z;
    }
    return rays;
}

// nrc_setup_kernels.cu:51-92
extern "C" void orc_nrc_accumulate(orc_frame* f, const GfxFrameParams* p) {
    orc_nrc_frame* n = nrcFrame(f);
    const uint32_t y0 = p->tileOriginY, y1 = p->tileRows ? std::min(f->H, p->tileOriginY + p->tileRows) : f->H;
    for (size_t i = (size_t)y0 * f->W; i < (size_t)y1 * f->W; ++i) {
        const uint32_t* t = &n->terminalInfo[4 * i];
        const float3 alpha(u2f(t[0]), u2f(t[1]), u2f(t[2]));
        const bool hasQuery = t[3] & 1u;
        const float3 directCont(n->frameContribution[3 * i], n->frameContribution[3 * i + 1], n->frameContribution[3 * i + 2]);
        float3 radiance(0.0f);
        if (hasQuery) {
            radiance = max3(float3(n->inferredRadiance[3 * i], n->inferredRadiance[3 * i + 1], n->inferredRadiance[3 * i + 2]), float3(0.0f));
            if (p->radianceScale > 0)
                radiance /= p->radianceScale;
            const float* q = &n->inferenceQuery[14 * i];
            radiance *= (float3(q[8], q[9], q[10]) + float3(q[11], q[12], q[13]));
        }
        const float3 indirectCont = alpha * radiance;
        const float3 contribution = directCont + indirectCont;
        float3 prevColorResult(0.0f);
        if (p->numAccumFrames > 0)
            prevColorResult = float3(f->beauty[i].x, f->beauty[i].y, f->beauty[i].z);
        const float curWeight = 1.0f / (1 + p->numAccumFrames);
        const float3 colorResult = (1 - curWeight) * prevColorResult + curWeight * contribution;
        f->beauty[i] = F4{ colorResult.x, colorResult.y, colorResult.z, 1.0f };
    }
}

// nrc_setup_kernels.cu:94-138
extern "C" void orc_nrc_propagate(orc_frame* f, const GfxFrameParams* p) {
    orc_nrc_frame* n = nrcFrame(f);
    const size_t offset = (size_t)f->W * f->H;
    for (uint32_t i = 0; i < n->maxNumTrainingSuffixes; ++i) {
        const uint32_t ti = n->suffixTerminal[i];
        const uint32_t prev = ti & 0x7FFFFFu;
        if (prev == kInvalidVertexDataIndex)
            continue;
        float3 contribution(0.0f);
        if ((ti >> 23) & 1u) {
            const float* r = &n->inferredRadiance[3 * (offset + i)];
            contribution = max3(float3(r[0], r[1], r[2]), float3(0.0f));
            if (p->radianceScale > 0)
                contribution /= p->radianceScale;
            const float* q = &n->inferenceQuery[14 * (offset + i)];
            contribution *= (float3(q[8], q[9], q[10]) + float3(q[11], q[12], q[13]));
        }
        uint32_t last = prev;
        while (last != kInvalidVertexDataIndex) {
            const uint32_t* vi = &n->trainVertexInfo[4 * (size_t)last];
            float* tgt = &n->trainTarget[0][3 * (size_t)last];
            const float3 indirectCont = float3(u2f(vi[0]), u2f(vi[1]), u2f(vi[2])) * contribution;
            contribution = float3(tgt[0], tgt[1], tgt[2]) + indirectCont;
            const float* q = &n->trainQuery[0][14 * (size_t)last];
            const float3 refFactor = float3(q[8], q[9], q[10]) + float3(q[11], q[12], q[13]);
            // safeDivide(RGB, RGB) (basic_types.h:5224-5229)
            tgt[0] = refFactor.x != 0 ? contribution.x / refFactor.x : 0.0f;
            tgt[1] = refFactor.y != 0 ? contribution.y / refFactor.y : 0.0f;
            tgt[2] = refFactor.z != 0 ? contribution.z / refFactor.z : 0.0f;
            last = vi[3] & 0x7FFFFFu;
        }
    }
    if (n->sharded()) { // all records on all ranks: every record is non-zero on one rank only, the integer sum keeps its bits
        n->shardSum(n->shardUser, reinterpret_cast<uint32_t*>(n->trainQuery[0].data()), n->trainQuery[0].size());
        n->shardSum(n->shardUser, reinterpret_cast<uint32_t*>(n->trainTarget[0].data()), n->trainTarget[0].size());
    }
}

// nrc_setup_kernels.cu:140-216.  targetAvg is accumulated in index order (the reference's float atomics have no
// defined order; the statistic only feeds the GUI).
extern "C" void orc_nrc_shuffle(orc_frame* f, const GfxFrameParams* p) {
    orc_nrc_frame* n = nrcFrame(f);
    const uint32_t bufIdx = p->bufferIndex & 1;
    const uint32_t numTrainingData = n->state[NRC_NUM_TRAINING_DATA + bufIdx];
    float avg[3] = { 0, 0, 0 };
    int32_t mn[3], mx[3];
    for (int c = 0; c < 3; ++c) {
        mn[c] = (int32_t)n->state[NRC_TARGET_MIN + 6 * bufIdx + c];
        mx[c] = (int32_t)n->state[NRC_TARGET_MAX + 6 * bufIdx + c];
    }
    for (uint32_t i = 0; i < kNumTrainingDataPerFrame; ++i) {
        if (numTrainingData > 0) {
            uint32_t &lcg = n->shufflers[i];
            lcg = (lcg * 1103515245u + 12345u) % (1u << 31);
            const uint32_t dstIdx = lcg % kNumTrainingDataPerFrame;
            // the path tracer never writes beyond trainBufferSize records
            const uint32_t srcIdx = i % std::min(numTrainingData, kTrainBufferSize);
            float query[14];
            std::memcpy(query, &n->trainQuery[0][14 * (size_t)srcIdx], sizeof(query));
            float tgt[3] = { n->trainTarget[0][3 * (size_t)srcIdx], n->trainTarget[0][3 * (size_t)srcIdx + 1], n->trainTarget[0][3 * (size_t)srcIdx + 2] };
            bool valid = true;
            for (int c = 0; c < 14; ++c)
                valid = valid && std::isfinite(query[c]);
            if (!valid)
                std::memset(query, 0, sizeof(query));
            if (!(std::isfinite(tgt[0]) && std::isfinite(tgt[1]) && std::isfinite(tgt[2])))
                tgt[0] = tgt[1] = tgt[2] = 0.0f;
            for (int c = 0; c < 3; ++c) {
                mn[c] = std::min(mn[c], floatToOrderedInt(tgt[c]));
                mx[c] = std::max(mx[c], floatToOrderedInt(tgt[c]));
                avg[c] += tgt[c] * (1.0f / kNumTrainingDataPerFrame);
                if (p->radianceScale > 0)
                    tgt[c] *= p->radianceScale;
                tgt[c] = std::fmin(tgt[c], 1e+6f);
            }
            std::memcpy(&n->trainQuery[1][14 * (size_t)dstIdx], query, sizeof(query));
            std::memcpy(&n->trainTarget[1][3 * (size_t)dstIdx], tgt, sizeof(tgt));
        }
        else {
            std::memset(&n->trainQuery[1][14 * (size_t)i], 0, 14 * sizeof(float));
            std::memset(&n->trainTarget[1][3 * (size_t)i], 0, 3 * sizeof(float));
        }
    }
    for (int c = 0; c < 3; ++c) {
        n->state[NRC_TARGET_MIN + 6 * bufIdx + c] = (uint32_t)mn[c];
        n->state[NRC_TARGET_MAX + 6 * bufIdx + c] = (uint32_t)mx[c];
        n->state[NRC_TARGET_AVG + 3 * bufIdx + c] = f2u(u2f(n->state[NRC_TARGET_AVG + 3 * bufIdx + c]) + avg[c]);
    }
}
