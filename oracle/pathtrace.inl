// oracle/pathtrace.inl — TEST INFRASTRUCTURE (CPU oracle), included by render.cpp.
//
// Per-pixel CPU restatement of the unidirectional path tracer of the reference:
//   performNextEventEstimation     path_tracing/gpu_kernels/optix_pathtracing_kernels.cu:18-71
//   pathTrace_rayGen_generic       path_tracing/gpu_kernels/optix_pathtracing_kernels.cu:73-216
//   pathTrace_closestHit_generic   path_tracing/gpu_kernels/optix_pathtracing_kernels.cu:218-300
//   pathTraceBaseline miss         path_tracing/gpu_kernels/optix_pathtracing_kernels.cu:310-341 (environment light, MIS)
//   computeSurfacePoint<true,false>path_tracing/path_tracing_shared.h:484-580 (closest hit, hypothetical area pdf)
//   computeSurfacePoint            path_tracing/path_tracing_shared.h:582-621 (first hit from GBuffer0)
// with the compile-time switches of :12-16: useSolidAngleSampling = false, implicit + explicit light
// sampling, MIS (power heuristic).  optixTrace -> the restated bvh::traverse (canonical tie-break);
// optixTransform{Point,Vector,Normal}FromObjectToWorldSpace -> the instance's transform / normalMatrix
// (OptiX internals, "parity unpinned" like every traversal result).  No bump mapping (1x1 textures).
// Environment light (envlight.h) when the scene has a map and params->enableEnvLight is set.
//
// RNG draw order per path vertex (PCG32, one state per pixel):
//   NEE: uLight, u0, u1      BSDF sample: uDir0, uDir1      (first hit)
//   then per closest hit: [RR u]  NEE(uLight,u0,u1)  BSDF(uDir0,uDir1)
// Function-call arguments `rng.getFloat0cTo1o(), rng.getFloat0cTo1o()` are taken left to right (the
// convention fixed in SURVEY.md Appendix A.3 for the ReSTIR candidates).

struct SurfacePoint {
    float3 positionInWorld, shadingNormalInWorld, texCoord0DirInWorld, geometricNormalInWorld;
    float2 texCoord;
    float hypAreaPDensity;
};

// path_tracing_shared.h:582-621
static void computeSurfacePointFromGBuffer(const orc_scene* s, const InstData &inst, const MeshData &mesh,
                                           uint32_t primIndex, float bcB, float bcC, SurfacePoint* sp) {
    const uint32_t i0 = mesh.triangles[3 * primIndex], i1 = mesh.triangles[3 * primIndex + 1], i2 = mesh.triangles[3 * primIndex + 2];
    const float3 pA = ld3(mesh.positions, i0), pB = ld3(mesh.positions, i1), pC = ld3(mesh.positions, i2);
    const float bcA = 1 - (bcB + bcC);
    const Affine xfm = affine(inst.desc.transform);
    const Mat3 normalMatrix = mat3(inst.desc.normalMatrix);

    const float3 positionInObj = bcA * pA + bcB * pB + bcC * pC;
    sp->positionInWorld = xfm.point(positionInObj);
    sp->geometricNormalInWorld = normalize(normalMatrix.mul(cross(pB - pA, pC - pA)));
    const float3 shadingNormalInObj = bcA * ld3(mesh.normals, i0) + bcB * ld3(mesh.normals, i1) + bcC * ld3(mesh.normals, i2);
    const float3 texCoord0DirInObj = bcA * ld3(mesh.tangents, i0) + bcB * ld3(mesh.tangents, i1) + bcC * ld3(mesh.tangents, i2);
    sp->texCoord = bcA * ld2(mesh.texcoords, i0) + bcB * ld2(mesh.texcoords, i1) + bcC * ld2(mesh.texcoords, i2);

    sp->shadingNormalInWorld = normalize(normalMatrix.mul(shadingNormalInObj));
    sp->texCoord0DirInWorld = xfm.vector(texCoord0DirInObj);
    sp->texCoord0DirInWorld = normalize(
        sp->texCoord0DirInWorld - dot(sp->shadingNormalInWorld, sp->texCoord0DirInWorld) * sp->shadingNormalInWorld);
    if (!allFinite(sp->shadingNormalInWorld)) {
        sp->geometricNormalInWorld = float3(0, 0, 1);
        sp->shadingNormalInWorld = float3(0, 0, 1);
        sp->texCoord0DirInWorld = float3(1, 0, 0);
    }
    if (!allFinite(sp->texCoord0DirInWorld)) {
        float3 bitangent;
        makeCoordinateSystem(sp->shadingNormalInWorld, &sp->texCoord0DirInWorld, &bitangent);
    }
    sp->hypAreaPDensity = 0.0f;
    (void)s;
}

// path_tracing_shared.h:484-580 with computeHypotheticalAreaPDensity = true, useSolidAngleSampling = false
static void computeSurfacePointAtHit(const orc_scene* s, const GfxFrameParams* p, const InstData &inst, const MeshData &mesh,
                                     uint32_t primIndex, float bcB, float bcC, SurfacePoint* sp) {
    const uint32_t i0 = mesh.triangles[3 * primIndex], i1 = mesh.triangles[3 * primIndex + 1], i2 = mesh.triangles[3 * primIndex + 2];
    const Affine xfm = affine(inst.desc.transform);
    const Mat3 normalMatrix = mat3(inst.desc.normalMatrix);
    const float3 pA = xfm.point(ld3(mesh.positions, i0));
    const float3 pB = xfm.point(ld3(mesh.positions, i1));
    const float3 pC = xfm.point(ld3(mesh.positions, i2));
    const float bcA = 1 - (bcB + bcC);

    sp->positionInWorld = bcA * pA + bcB * pB + bcC * pC;
    const float3 shadingNormalInObj = bcA * ld3(mesh.normals, i0) + bcB * ld3(mesh.normals, i1) + bcC * ld3(mesh.normals, i2);
    const float3 texCoord0DirInObj = bcA * ld3(mesh.tangents, i0) + bcB * ld3(mesh.tangents, i1) + bcC * ld3(mesh.tangents, i2);
    sp->texCoord = bcA * ld2(mesh.texcoords, i0) + bcB * ld2(mesh.texcoords, i1) + bcC * ld2(mesh.texcoords, i2);

    sp->geometricNormalInWorld = cross(pB - pA, pC - pA);
    const float area = 0.5f * length(sp->geometricNormalInWorld);
    sp->geometricNormalInWorld = sp->geometricNormalInWorld / (2 * area);

    sp->shadingNormalInWorld = normalize(normalMatrix.mul(shadingNormalInObj));
    sp->texCoord0DirInWorld = normalize(xfm.vector(texCoord0DirInObj));
    if (!allFinite(sp->shadingNormalInWorld)) {
        sp->shadingNormalInWorld = float3(0, 0, 1);
        sp->texCoord0DirInWorld = float3(1, 0, 0);
    }
    if (!allFinite(sp->texCoord0DirInWorld)) {
        float3 bitangent;
        makeCoordinateSystem(sp->shadingNormalInWorld, &sp->texCoord0DirInWorld, &bitangent);
    }

    float lightProb = 1.0f;
    if (useEnvLight(s, p)) // path_tracing_shared.h:540-541
        lightProb *= (1 - kProbToSampleEnvLight);
    const float instImportance = inst.geomIntegral;
    lightProb *= (pow2(inst.desc.uniformScale) * instImportance) / s->instIntegral;
    lightProb *= mesh.primIntegral / instImportance;
    if (!std::isfinite(lightProb)) {
        sp->hypAreaPDensity = 0.0f;
        return;
    }
    // DiscreteDistribution1D::evaluatePMF (common_shared.h:248-253)
    lightProb *= (mesh.primWeights.empty() || mesh.primIntegral == 0.0f) ? 0.0f : mesh.primWeights[primIndex] / mesh.primIntegral;
    sp->hypAreaPDensity = lightProb / area;
}

// the miss programs of the path tracers (optix_pathtracing_kernels.cu:310-341; the NRC one :625-650 has the same body):
// luminance * misWeight of the environment along rayDir, zero without an environment light
static float3 evaluateEnvLightOnMiss(const orc_scene* s, const GfxFrameParams* p, const float3 &rayDirIn, float prevDirPDensity,
                                     bool nrcVariant = false) {
    if (!useEnvLight(s, p))
        return float3(0.0f);
    const float3 rayDir = normalize(rayDirIn);
    float posPhi, theta;
    toPolarYUp(rayDir, &posPhi, &theta);
    float phi = posPhi + p->envLightRotation;
    phi = phi - std::floor(phi / (2 * kPi)) * 2 * kPi;
    const float2 texCoord(phi / (2 * kPi), theta / kPi);
    const float3 luminance = p->envLightPowerCoeff * s->env.fetch(texCoord.x, texCoord.y);
    const float uvPDF = s->env.evaluatePDF(texCoord.x, texCoord.y);
    const float hypAreaPDensity = uvPDF / (2 * kPi * kPi * dm_sin(theta));
    // the NRC app's miss program multiplies by probToSampleEnvLight unconditionally (neural_radiance_caching/.../:646)
    const float lightPDensity = (nrcVariant || s->instIntegral > 0.0f ? kProbToSampleEnvLight : 1.0f) * hypAreaPDensity;
    const float bsdfPDensity = prevDirPDensity;
    const float misWeight = pow2(bsdfPDensity) / (pow2(bsdfPDensity) + pow2(lightPDensity));
    return luminance * misWeight;
}

struct PathTraceCounters {
    uint64_t closestRays = 0, visibilityRays = 0;
};

// optix_pathtracing_kernels.cu:18-71
static float3 performNextEventEstimation(const orc_scene* s, const GfxFrameParams* p, const float3 &shadingPoint, const float3 &vOutLocal,
                                         const ReferenceFrame &shadingFrame, const BSDF &bsdf, PCG32RNG &rng,
                                         PathTraceCounters* counters) {
    float3 ret(0.0f);
    float uLight = rng.getFloat0cTo1o();
    bool selectEnvLight = false;
    float probToSampleCurLightType = 1.0f;
    if (useEnvLight(s, p)) { // :27-42
        if (s->instIntegral > 0.0f) {
            if (uLight < kProbToSampleEnvLight) {
                probToSampleCurLightType = kProbToSampleEnvLight;
                uLight /= probToSampleCurLightType;
                selectEnvLight = true;
            }
            else {
                probToSampleCurLightType = 1.0f - kProbToSampleEnvLight;
                uLight = (uLight - kProbToSampleEnvLight) / probToSampleCurLightType;
            }
        }
        else {
            selectEnvLight = true;
        }
    }
    LightSample lightSample;
    float areaPDensity = 0.0f;
    const float u0 = rng.getFloat0cTo1o();
    const float u1 = rng.getFloat0cTo1o();
    sampleLight(s, p, uLight, selectEnvLight, u0, u1, &lightSample, &areaPDensity);
    areaPDensity *= probToSampleCurLightType;
    if (areaPDensity > 0.0f) {
        float3 shadowRay = lightSample.atInfinity ? lightSample.position : (lightSample.position - shadingPoint);
        const float dist2 = sqLength(shadowRay);
        shadowRay /= std::sqrt(dist2);
        const float3 vInLocal = shadingFrame.toLocal(shadowRay);
        const float lpCos = std::fabs(dot(shadowRay, lightSample.normal));
        float bsdfPDensity = bsdf.evaluatePDF(vOutLocal, vInLocal) * lpCos / dist2;
        if (!std::isfinite(bsdfPDensity))
            bsdfPDensity = 0.0f;
        const float lightPDensity = areaPDensity;
        const float misWeight = pow2(lightPDensity) / (pow2(bsdfPDensity) + pow2(lightPDensity));
        ++counters->visibilityRays;
        ret = performDirectLighting<true>(s, shadingPoint, vOutLocal, shadingFrame, bsdf, lightSample) * (misWeight / areaPDensity);
    }
    return ret;
}

static void pathTracePixel(orc_frame* f, const GfxFrameParams* p, const Camera &camera, uint32_t x, uint32_t y,
                           PathTraceCounters* counters) {
    const orc_scene* s = f->scene;
    const size_t pix = (size_t)y * f->W + x;
    const uint32_t bufIdx = p->bufferIndex & 1;
    const GB0 gb0 = f->gb0[bufIdx][pix];
    const float bcB = decodeBarycentric((uint16_t)(gb0.qbc & 0xFFFFu));
    const float bcC = decodeBarycentric((uint16_t)(gb0.qbc >> 16));
    const uint32_t maxPathLength = p->maxPathLength ? p->maxPathLength : 5u; // 0 = the hosts' default (path_tracing_main.cpp:1519)

    float3 contribution(0.001f, 0.001f, 0.001f);
    if (gb0.instSlot != 0xFFFFFFFFu) {
        const InstData* inst = &s->instances[gb0.instSlot];
        const MeshData* mesh = &s->meshes[gb0.geomInstSlot];
        SurfacePoint sp;
        computeSurfacePointFromGBuffer(s, *inst, *mesh, gb0.primIndex, bcB, bcC, &sp);

        float3 alpha(1.0f);
        const float initImportance = sRGB_calcLuminance(alpha);
        PCG32RNG rng{ f->rng[pix] };

        // shading on the first hit (:106-146)
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
            contribution += alpha * performNextEventEstimation(s, p, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, counters);

            float3 vInLocal;
            const float uDir0 = rng.getFloat0cTo1o();
            const float uDir1 = rng.getFloat0cTo1o();
            alpha *= bsdf.sampleThroughput(vOutLocal, uDir0, uDir1, &vInLocal, &dirPDensity);
            vIn = shadingFrame.fromLocal(vInLocal);
        }

        // path extension loop (:148-194) with the closest-hit program (:218-300) inlined
        float prevDirPDensity = dirPDensity;
        uint32_t pathLength = 1;
        float3 rayOrg = positionInWorld;
        float3 rayDir = vIn;
        while (true) {
            const bool isValidSampling = prevDirPDensity > 0.0f && std::isfinite(prevDirPDensity);
            if (!isValidSampling)
                break;
            ++pathLength;
            const bool maxLengthTerminate = pathLength >= maxPathLength;

            ++counters->closestRays;
            const HitObject hit = traverseCanonical(s->bvh, rayOrg, rayDir, 0.0f, std::numeric_limits<float>::max());
            if (hit.primIndex == UINT32_MAX) { // miss program (:310-341): implicit environment light sampling with MIS
                contribution += alpha * evaluateEnvLightOnMiss(s, p, rayDir, prevDirPDensity);
                break;
            }

            inst = &s->instances[s->geomToInst[hit.geomIndex]];
            mesh = &s->meshes[s->geomToMesh[hit.geomIndex]];
            computeSurfacePointAtHit(s, p, *inst, *mesh, hit.primIndex, hit.bcB, hit.bcC, &sp);
            const GfxMaterialDesc &mat = s->materials[mesh->materialSlot];

            const float3 vOut = normalize(-rayDir);
            const float frontHit = dot(vOut, sp.geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
            const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
            positionInWorld = offsetRayOrigin(sp.positionInWorld, frontHit * sp.geometricNormalInWorld);
            const float3 vOutLocal = shadingFrame.toLocal(vOut);

            // implicit light sampling with MIS (:262-275)
            if (vOutLocal.z > 0 && mat.hasEmittance) {
                const float3 emittance(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
                const float dist2 = sqLength(positionInWorld - rayOrg); // sqDistance(rayOrigin, positionInWorld)
                const float lightPDensity = sp.hypAreaPDensity * dist2 / vOutLocal.z;
                const float bsdfPDensity = prevDirPDensity;
                const float misWeight = pow2(bsdfPDensity) / (pow2(bsdfPDensity) + pow2(lightPDensity));
                contribution += alpha * emittance * (misWeight / kPi);
            }

            // Russian roulette (:277-281)
            const float continueProb = std::fmin(sRGB_calcLuminance(alpha) / initImportance, 1.0f);
            if (rng.getFloat0cTo1o() >= continueProb || maxLengthTerminate)
                break;
            alpha /= continueProb;

            const BSDF bsdf = setupBsdf(s, mesh->materialSlot, sp.texCoord);
            contribution += alpha * performNextEventEstimation(s, p, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, counters);

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
    else if (useEnvLight(s, p)) { // :200-207: the environment seen directly; the miss program left (u, v) in the barycentrics
        contribution = p->envLightPowerCoeff * s->env.fetch(bcB, bcC);
    }

    float3 prevColorResult(0.0f);
    if (p->numAccumFrames > 0)
        prevColorResult = float3(f->beauty[pix].x, f->beauty[pix].y, f->beauty[pix].z);
    const float curWeight = 1.0f / (1 + p->numAccumFrames);
    const float3 colorResult = (1 - curWeight) * prevColorResult + curWeight * contribution;
    f->beauty[pix] = F4{ colorResult.x, colorResult.y, colorResult.z, 1.0f };
}

static uint64_t nrcPathTrace(orc_frame* f, const GfxFrameParams* p, int numThreads); // nrc_pathtrace.inl
static uint64_t regirPathTrace(orc_frame* f, const GfxFrameParams* p, int numThreads); // regir.inl

extern "C" uint64_t orc_pathtrace(orc_frame* f, const GfxFrameParams* p, int variant, int numThreads) {
    if (numThreads <= 0) numThreads = omp_get_max_threads();
    if (variant == GFX_PT_NRC)
        return nrcPathTrace(f, p, numThreads);
    if (variant == GFX_PT_REGIR)
        return regirPathTrace(f, p, numThreads);
    const Camera camera = makeCamera(p->camera);
    const uint32_t W = f->W, H = f->H;
    const uint32_t y0 = p->tileOriginY, y1 = p->tileRows ? std::min(H, p->tileOriginY + p->tileRows) : H;
    uint64_t rays = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(numThreads) reduction(+ : rays)
    for (int64_t yy = y0; yy < (int64_t)y1; ++yy) {
        PathTraceCounters counters;
        for (uint32_t x = 0; x < W; ++x)
            pathTracePixel(f, p, camera, x, (uint32_t)yy, &counters);
        rays += counters.closestRays + counters.visibilityRays;
    }
    return rays;
}
