// lighting.cuh — light sampling and direct-lighting helpers shared by the ReSTIR, ReGIR and path-tracing kernels.
//
// sampleLight<false> (restir_di/restir_di_shared.h:320-516; identical copies live in path_tracing_shared.h:220-482,
// regir_shared.h, neural_radiance_caching_shared.h, svgf_shared.h), performDirectLighting (:518-557),
// evaluateVisibility (:559-582).  Visibility rays run through the inline any-hit traversal of traverse.cuh
// (fused kernels) or are queued for trace.cu's persistent kernel (wavefront kernels).
#pragma once
#include "traverse.cuh"
#include "shading.cuh"

namespace gfx {

struct LightSample { // restir_di_shared.h:89-96
    f3 emittance, position, normal;
    uint32_t atInfinity;
};
GFX_D LightSample emptyLightSample() {
    LightSample s;
    s.emittance = f3(0.0f);
    s.position = f3(0.0f);
    s.normal = f3(0.0f);
    s.atInfinity = 0;
    return s;
}

GFX_D float convertToWeight(const f3 &c) { return (c.x + c.y + c.z) / 3; } // restir_di_shared.h:82-85

// last index with cdf[idx] <= u: the power-of-two stepping search of
// DiscreteDistribution1DTemplate::sample (common_shared.h:226-232)
GFX_D uint32_t searchCdf(const float* __restrict__ cdf, uint32_t numValues, float u) {
    int idx = 0;
    for (int d = (int)(nextPowerOf2(numValues) >> 1); d >= 1; d >>= 1) {
        if (idx + d >= (int)numValues)
            continue;
        if (__ldg(cdf + idx + d) <= u)
            idx += d;
    }
    return (uint32_t)idx;
}
// the same search seeded by a guide table (scene.cuh): exact, 1-3 loads instead of log2(n)
GFX_D uint32_t guidedSearchCdf(const float* __restrict__ cdf, const uint32_t* __restrict__ guide, uint32_t guideSize,
                               float u01, float u) {
    const uint32_t b = min(dm_f2uint(u01 * (float)guideSize), guideSize - 1);
    uint32_t idx = __ldg(guide + b);
    const uint32_t hi = __ldg(guide + b + 1);
    while (idx < hi && __ldg(cdf + idx + 1) <= u)
        ++idx;
    return idx;
}
GFX_D float remapCdf(const float* __restrict__ cdf, uint32_t numValues, float integral, uint32_t idx, float u) {
    // common_shared.h:235-241
    const float lCDF = __ldg(cdf + idx);
    float rCDF = integral;
    if (idx < numValues - 1)
        rCDF = __ldg(cdf + idx + 1);
    return (u - lCDF) / (rCDF - lCDF);
}

// The light-sampling fields of an instance and its normal matrix, as scalar loads or (GFX_WIDE_TABLE_LOADS, scene.cuh) as
// wide ones; the values are the same either way.
struct InstanceSampling {
    uint32_t firstMeshSlot, numMeshSlots;
    float geomIntegral;
    uint32_t geomBase;
};
GFX_D InstanceSampling loadInstanceSampling(const DevInstance* inst) {
    InstanceSampling r;
#if GFX_WIDE_TABLE_LOADS
    static_assert(offsetof(DevInstance, firstMeshSlot) % 16 == 0, "the sampling fields must start a 16-byte line");
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(&inst->firstMeshSlot));
    r.firstMeshSlot = q.x;
    r.numMeshSlots = q.y;
    r.geomIntegral = __uint_as_float(q.z);
    r.geomBase = q.w;
#else
    r.firstMeshSlot = inst->firstMeshSlot;
    r.numMeshSlots = inst->numMeshSlots;
    r.geomIntegral = inst->geomIntegral;
    r.geomBase = inst->geomBase;
#endif
    return r;
}
GFX_D f3 applyNormalMatrix(const DevInstance* inst, const f3 &n) {
#if GFX_WIDE_TABLE_LOADS
    static_assert(offsetof(DevInstance, normalMatrix) % 16 == 0, "the normal matrix must start a 16-byte line");
    const float4 a = __ldg(reinterpret_cast<const float4*>(inst->normalMatrix));
    const float4 b = __ldg(reinterpret_cast<const float4*>(inst->normalMatrix) + 1);
    const float m[9] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, __ldg(inst->normalMatrix + 8) };
    return mul3x3(m, n);
#else
    return mul3x3(inst->normalMatrix, n);
#endif
}

// The three DiscreteDistribution1D::sample calls of sampleLight (instance, geometry instance, primitive) spelled out;
// weights[idx] / integral comes from the pre-divided prob tables.  Returns false on sampleLight's probability-0 early outs.
struct LightTrianglePick {
    const DevInstance* inst;
    uint32_t lightTri;   // index into the lightTris table (lights.cu)
    float lightProb;
};
GFX_D bool pickLightTriangle(const DevScene &s, float ul, LightTrianglePick* pick) {
    float lightProb = 1.0f;

    // instance
    const float instIntegral = __ldg(s.instIntegral);
    float u = ul * instIntegral;
    const uint32_t instSlot = guidedSearchCdf(s.instCdf, s.instGuide, kInstGuideSize, ul, u);
    const float uGeomInst = remapCdf(s.instCdf, s.numInstances, instIntegral, instSlot, u);
    const float instProb = __ldg(s.instProb + instSlot);
    lightProb *= instProb;
    const DevInstance* inst = s.instances + instSlot;
    if (instProb == 0.0f)
        return false;

    // geometry instance
#if GFX_WIDE_TABLE_LOADS
    const InstanceSampling is = loadInstanceSampling(inst);
    const uint32_t firstMeshSlot = is.firstMeshSlot, numMeshSlots = is.numMeshSlots;
    const float geomIntegral = is.geomIntegral;
    const uint32_t geomBase = is.geomBase;
#else
    const uint32_t firstMeshSlot = inst->firstMeshSlot, numMeshSlots = inst->numMeshSlots;
    const float geomIntegral = inst->geomIntegral;
#endif
    u = uGeomInst * geomIntegral;
    const uint32_t geomInstIndexInInst = searchCdf(s.geomCdf + firstMeshSlot, numMeshSlots, u);
    const float uPrim = remapCdf(s.geomCdf + firstMeshSlot, numMeshSlots, geomIntegral, geomInstIndexInInst, u);
    const float geomInstProb = __ldg(s.geomProb + firstMeshSlot + geomInstIndexInInst);
    const uint32_t geomInstSlot = __ldg(s.instanceMeshSlots + firstMeshSlot + geomInstIndexInInst);
    lightProb *= geomInstProb;
    if (geomInstProb == 0.0f)
        return false;

    // primitive
    const DevMesh* mesh = s.meshes + geomInstSlot;
    const uint32_t triBase = mesh->triBase;
    u = uPrim * mesh->primIntegral;
    const uint32_t primIndex = guidedSearchCdf(s.primCdf + triBase, s.primGuide + (size_t)geomInstSlot * (kPrimGuideSize + 1),
                                               kPrimGuideSize, uPrim, u);
    const float primProb = __ldg(s.primProb + triBase + primIndex);
    lightProb *= primProb;

    pick->inst = inst;
#if GFX_WIDE_TABLE_LOADS
    pick->lightTri = __ldg(s.lightTriBase + geomBase + geomInstIndexInInst) + primIndex;
#else
    pick->lightTri = __ldg(s.lightTriBase + inst->geomBase + geomInstIndexInInst) + primIndex;
#endif
    pick->lightProb = lightProb;
    return true;
}

// A Low-Distortion Map Between Triangle and Square (restir_di_shared.h:485-498)
GFX_D void squareToTriangle(float u0, float u1, float* bcA, float* bcB, float* bcC) {
    float a = 0.5f * u0;
    float b = 0.5f * u1;
    const float offset = b - a;
    if (offset > 0)
        b += offset;
    else
        a -= offset;
    *bcA = a;
    *bcB = b;
    *bcC = 1 - (a + b);
}

GFX_D void sampleLight(const DevScene &s, float ul, float u0, float u1, LightSample* lightSample, float* areaPDensity) {
    // restir_di_shared.h:320-516 with sampleEnvLight = false, useSolidAngleSampling = false; the triangle operands come
    // from the lightTris table (lights.cu), produced by the reference's own expressions, so the result is bit-identical.
    LightTrianglePick pick;
    if (!pickLightTriangle(s, ul, &pick)) {
        *areaPDensity = 0.0f;
        return;
    }
    const float4* e = s.lightTris + kLightTriStride * (size_t)pick.lightTri;
    const float4 e0 = __ldg(e + 0), e1 = __ldg(e + 1), e2 = __ldg(e + 2), e3 = __ldg(e + 3), e4 = __ldg(e + 4), e5 = __ldg(e + 5);
    const f3 pA(e0.x, e0.y, e0.z), pB(e1.x, e1.y, e1.z), pC(e2.x, e2.y, e2.z);
    const f3 nA(e1.w, e2.w, e3.x), nB(e3.y, e3.z, e3.w), nC(e4.x, e4.y, e4.z);
    float bcA, bcB, bcC;
    squareToTriangle(u0, u1, &bcA, &bcB, &bcC);

    const float recArea = e0.w;
    *areaPDensity = pick.lightProb * recArea;

    lightSample->position = bcA * pA + bcB * pB + bcC * pC;
    lightSample->atInfinity = 0;
    lightSample->normal = bcA * nA + bcB * nB + bcC * nC;
    lightSample->normal = normalize(applyNormalMatrix(pick.inst, lightSample->normal));
    lightSample->emittance = f3(e5.x, e5.y, e5.z);
}

// sampleLight for the RIS candidate loop, fetching the 96-byte light triangle in three steps and stopping as soon as
// the candidate is certain to contribute RGB(0) to this shading point.  ncu: the candidate kernel runs the L1 data pipe
// at 84 % of its wavefront rate (535 M wavefronts for 77 M load instructions - every lane reads a different light), and
// the six 16-byte fetches of the light triangle are three quarters of them; ~70 % of the candidates are dark.
//   step 1: the three vertices -> sample position.  Below the shading horizon (the BRDFs return RGB(0) when
//           vGiven.z * vSampled.z <= 0)?  -> dark, 3 fetches and the instance's normal matrix saved;
//   step 2: the vertex normals -> un-normalised light normal.  Facing away (lpCos <= 0)? -> dark, 1 fetch saved;
//   step 3: the emittance, the normalisation: the full sample.
// Both tests run on un-normalised vectors with a relative margin of 1e-3 on the cosine - three orders of magnitude above
// the rounding of the exact evaluation - so borderline samples go on to the exact path and no decision ever differs from
// performDirectLighting's; NaNs and a zero distance fail the comparisons and also go on.  Returns true for "dark" (then
// *areaPDensity is the sample's density and is positive); otherwise the outputs are sampleLight's, bit for bit.
GFX_D bool sampleLightUnlessDark(const DevScene &s, float ul, float u0, float u1, const f3 &shadingPoint,
                                 const f3 &shadingNormal, float vOutLocalZ, LightSample* lightSample, float* areaPDensity) {
    LightTrianglePick pick;
    if (!pickLightTriangle(s, ul, &pick)) {
        *areaPDensity = 0.0f;
        return false;
    }
    const float4* e = s.lightTris + kLightTriStride * (size_t)pick.lightTri;
#if GFX_LIGHT_CULL_SPHERES
    {   // step 0: the whole triangle lies below the shading horizon if its bounding sphere does - with twice the margin of the
        // per-sample test below, so that every sample point of the triangle would pass that test too:
        //   -dot(q - p, n) sgn >= t - r |n|  and  |q - p| <= |c - p| + r   for every q in the sphere
        const float4 sphere = __ldg(e + 6);
        const f3 dc = f3(sphere.x, sphere.y, sphere.z) - shadingPoint;
        const float r = sphere.w;
        const float t = -(dot(dc, shadingNormal) * vOutLocalZ);          // > 0: centre on the far side of the surface
        const float margin = t - 1.001f * r * fabsf(vOutLocalZ);
        if (r >= 0.0f && pick.lightProb > 0.0f && margin > 0.0f &&
            margin * margin > 8e-6f * (sqLength(dc) + r * r) * (vOutLocalZ * vOutLocalZ)) {
            *areaPDensity = pick.lightProb; // any positive number: the caller only asks whether the density is positive
            return true;
        }
    }
#endif
    const float4 e0 = __ldg(e + 0), e1 = __ldg(e + 1), e2 = __ldg(e + 2);
    const f3 pA(e0.x, e0.y, e0.z), pB(e1.x, e1.y, e1.z), pC(e2.x, e2.y, e2.z);
    float bcA, bcB, bcC;
    squareToTriangle(u0, u1, &bcA, &bcB, &bcC);
    const float recArea = e0.w;
    const float density = pick.lightProb * recArea;
    *areaPDensity = density;
    const f3 position = bcA * pA + bcB * pB + bcC * pC;

    const float k = 1e-6f; // (1e-3)^2
    const f3 d = position - shadingPoint;
    const float dd = sqLength(d);
    const float b = dot(d, shadingNormal) * vOutLocalZ; // < 0: light and viewer on opposite sides of the surface
    if (density > 0.0f && b < 0.0f && b * b > k * dd * (vOutLocalZ * vOutLocalZ))
        return true;

    const float4 e3 = __ldg(e + 3), e4 = __ldg(e + 4);
    const f3 nA(e1.w, e2.w, e3.x), nB(e3.y, e3.z, e3.w), nC(e4.x, e4.y, e4.z);
    f3 normal = bcA * nA + bcB * nB + bcC * nC;
    normal = applyNormalMatrix(pick.inst, normal);
    const float a = dot(d, normal); // > 0: the emitter faces away from the shading point
    if (density > 0.0f && a > 0.0f && a * a > k * dd * sqLength(normal))
        return true;

    const float4 e5 = __ldg(e + 5);
    lightSample->position = position;
    lightSample->atInfinity = 0;
    lightSample->normal = normalize(normal);
    lightSample->emittance = f3(e5.x, e5.y, e5.z);
    return false;
}

GFX_D bool traceVisibility(const DevScene &s, const f3 &org, const f3 &dir, float tmax) {
    // ray statistics: one atomic per warp per call site (lanes currently active here)
    const uint32_t active = __activemask();
    if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == (uint32_t)(__ffs(active) - 1))
        atomicAdd(s.rayCounter, (unsigned long long)__popc(active));
    const Hit h = traverseBvh<true>(s.bvh, org, dir, 0.0f, tmax);
    return h.storageIndex == 0xFFFFFFFFu;
}

GFX_D bool evaluateVisibility(const DevScene &s, const f3 &shadingPoint, const LightSample &ls) { // :559-582
    f3 shadowRayDir = ls.atInfinity ? ls.position : (ls.position - shadingPoint);
    const float dist2 = sqLength(shadowRayDir);
    float dist = sqrtf(dist2);
    shadowRayDir /= dist;
    if (ls.atInfinity)
        dist = 1e+10f;
    return traceVisibility(s, shadingPoint, shadowRayDir, dist * 0.9999f);
}

template <bool withVisibility>
GFX_D f3 performDirectLighting(const DevScene &s, const f3 &shadingPoint, const f3 &vOutLocal,
                               const ReferenceFrame &shadingFrame, const BSDF &bsdf, const LightSample &ls) { // :518-557
    f3 shadowRayDir = ls.atInfinity ? ls.position : (ls.position - shadingPoint);
    const float dist2 = sqLength(shadowRayDir);
    float dist = sqrtf(dist2);
    shadowRayDir /= dist;
    const f3 shadowRayDirLocal = shadingFrame.toLocal(shadowRayDir);

    const float lpCos = dot(-shadowRayDir, ls.normal);
    const float spCos = shadowRayDirLocal.z;

    float visibility = 1.0f;
    if (withVisibility) {
        if (ls.atInfinity)
            dist = 1e+10f;
        if (!traceVisibility(s, shadingPoint, shadowRayDir, dist * 0.9999f))
            visibility = 0.0f;
    }

    if (visibility > 0 && lpCos > 0) {
        const f3 Le = ls.emittance / kPi;
        const f3 fsValue = bsdf.evaluate(vOutLocal, shadowRayDirLocal);
        const float G = lpCos * fabsf(spCos) / dist2;
        return fsValue * Le * G;
    }
    return f3(0.0f);
}

GFX_D BSDF setupBsdf(const DevScene &s, uint32_t matSlot) {
    const GfxMaterialDesc* m = s.materials + matSlot;
    BSDF b;
    b.setup(m->bsdfType, m->p0, m->p1, m->p2);
    return b;
}

} // namespace gfx
