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

// sm_100's 256-bit read-only load (LDG.E.256): one L1 tag lookup for 32 bytes.  volatile: ptxas would otherwise hoist the
// later stages' loads of sampleLightUnlessDark above the dark tests (it did, ncu round 2: all nine record loads were issued by
// every survivor of the sphere test), which is exactly the traffic the staging is there to avoid.
struct F8 { float4 lo, hi; };
GFX_D F8 ldg256(const float4* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
}
GFX_D float ldg32v(const float* p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// the normal matrix of an instance out of the compact table (scene.cuh): one 256-bit load + one scalar
GFX_D f3 applyNormalMatrix(const DevScene &s, uint32_t instSlot, const f3 &n) {
    const float4* mp = s.normalMats + (size_t)kNormalMatStride * instSlot;
    const F8 a = ldg256(mp);
    const float m[9] = { a.lo.x, a.lo.y, a.lo.z, a.lo.w, a.hi.x, a.hi.y, a.hi.z, a.hi.w, ldg32v(reinterpret_cast<const float*>(mp + 2)) };
    return mul3x3(m, n);
}

// The three DiscreteDistribution1D::sample calls of sampleLight (instance, geometry instance, primitive;
// restir_di_shared.h:356-409, common_shared.h:209-246) spelled out; weights[idx] / integral comes from the pre-divided prob
// tables.  This is the DEFINITION of the light pick; the kernels use its flattened form (pickLightTriangle below), which
// lights.cu derives from it.  Returns a key: the index of the light record, or kPickNone | position for sampleLight's
// probability-0 early outs (distinct positions give distinct keys, so that equal keys at two values of ul mean "same piece").
GFX_D uint32_t chainPickLightTriangle(const DevScene &s, float ul, float* lightProbOut) {
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
        return kPickNone | 0x20000000u | instSlot;

    // geometry instance
    const uint32_t firstMeshSlot = inst->firstMeshSlot, numMeshSlots = inst->numMeshSlots;
    const float geomIntegral = inst->geomIntegral;
    u = uGeomInst * geomIntegral;
    const uint32_t geomInstIndexInInst = searchCdf(s.geomCdf + firstMeshSlot, numMeshSlots, u);
    const float uPrim = remapCdf(s.geomCdf + firstMeshSlot, numMeshSlots, geomIntegral, geomInstIndexInInst, u);
    const float geomInstProb = __ldg(s.geomProb + firstMeshSlot + geomInstIndexInInst);
    const uint32_t geomInstSlot = __ldg(s.instanceMeshSlots + firstMeshSlot + geomInstIndexInInst);
    lightProb *= geomInstProb;
    if (geomInstProb == 0.0f)
        return kPickNone | (firstMeshSlot + geomInstIndexInInst);

    // primitive
    const DevMesh* mesh = s.meshes + geomInstSlot;
    const uint32_t triBase = mesh->triBase;
    u = uPrim * mesh->primIntegral;
    const uint32_t primIndex = guidedSearchCdf(s.primCdf + triBase, s.primGuide + (size_t)geomInstSlot * (kPrimGuideSize + 1),
                                               kPrimGuideSize, uPrim, u);
    const float primProb = __ldg(s.primProb + triBase + primIndex);
    lightProb *= primProb;

    *lightProbOut = lightProb;
    const uint32_t base = __ldg(s.lightTriBase + inst->geomBase + geomInstIndexInInst);
    if (base == 0xFFFFFFFFu) // a geometry without an emitter: only reachable when every weight is 0 (0 / 0 probabilities)
        return kPickNone | 0x10000000u | (firstMeshSlot + geomInstIndexInInst);
    return base + primIndex;
}

// The flattened pick (scene.cuh, lights.cu): the piece that contains ul - its key and the head of its light record (cull
// sphere, area density, instance slot), which for a pure bucket arrives with the guide entry itself.
struct LightPick {
    uint32_t key;      // light record index, or kPickNone set
    float4 sphere;     // centre, radius (< 0: do not cull)
    float density;     // lightProb * recArea
    uint32_t instSlot;
};
GFX_D F8 pickGuideFetch(const DevScene &s, float ul) { // the guide entry of ul's bucket (issue early, resolve later)
    const uint32_t b = min(dm_f2uint(ul * (float)kPickGuideSize), kPickGuideSize - 1); // exact: a power-of-two product, floored
    return ldg256(s.pickGuide + 2 * (size_t)b);
}
GFX_D LightPick pickLightFromGuide(const DevScene &s, float ul, const F8 &g);
GFX_D LightPick pickLight(const DevScene &s, float ul) {
#ifndef GFX_AB_CHAIN_PICK
    return pickLightFromGuide(s, ul, pickGuideFetch(s, ul));
#else
    LightPick r;
    // A/B (-DGFX_AB_CHAIN_PICK): the definition instead of its flattened form
    float unusedProb;
    r.key = chainPickLightTriangle(s, ul, &unusedProb);
    if (!(r.key & kPickNone)) {
        const F8 h0 = ldg256(s.lightTris + kLightTriStride * (size_t)r.key);
        r.sphere = h0.lo;
        r.density = h0.hi.x;
        r.instSlot = __float_as_uint(h0.hi.y);
    }
    return r;
#endif
}
GFX_D LightPick pickLightFromGuide(const DevScene &s, float ul, const F8 &g) {
    LightPick r;
    const uint32_t head = __float_as_uint(g.lo.x);
    if (head & kPickPure) {
        r.key = head & ~kPickPure;
        r.sphere = make_float4(g.lo.y, g.lo.z, g.lo.w, g.hi.x);
        r.density = g.hi.y;
        r.instSlot = __float_as_uint(g.hi.z);
        return r;
    }
    const uint32_t ulBits = __float_as_uint(ul);
    uint32_t idx = head;
    uint2 piece = __ldg(s.pickPieces + idx);
    for (;;) {
        const uint2 next = __ldg(s.pickPieces + idx + 1);
        if (next.x > ulBits)
            break;
        piece = next;
        ++idx;
    }
    r.key = piece.y;
    r.sphere = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
    r.density = 0.0f;
    r.instSlot = 0;
    if (!(r.key & kPickNone)) {
        const F8 h0 = ldg256(s.lightTris + kLightTriStride * (size_t)r.key);
        r.sphere = h0.lo;
        r.density = h0.hi.x;
        r.instSlot = __float_as_uint(h0.hi.y);
    }
    return r;
}
GFX_D uint32_t pickLightTriangle(const DevScene &s, float ul) { return pickLight(s, ul).key; }

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
    // from the light records (lights.cu), produced by the reference's own expressions, so the result is bit-identical.
    const LightPick pick = pickLight(s, ul);
    if (pick.key & kPickNone) {
        *areaPDensity = 0.0f;
        return;
    }
    const float4* e = s.lightTris + kLightTriStride * (size_t)pick.key;
    const F8 h1 = ldg256(e + 2), h2 = ldg256(e + 4), h3 = ldg256(e + 6);
    const f3 pA(h1.lo.x, h1.lo.y, h1.lo.z), pB(h1.lo.w, h1.hi.x, h1.hi.y), pC(h1.hi.z, h1.hi.w, h2.lo.x);
    const f3 nA(h2.lo.y, h2.lo.z, h2.lo.w), nB(h2.hi.x, h2.hi.y, h2.hi.z), nC(h2.hi.w, h3.lo.x, h3.lo.y);
    float bcA, bcB, bcC;
    squareToTriangle(u0, u1, &bcA, &bcB, &bcC);

    *areaPDensity = pick.density;

    lightSample->position = bcA * pA + bcB * pB + bcC * pC;
    lightSample->atInfinity = 0;
    lightSample->normal = bcA * nA + bcB * nB + bcC * nC;
    lightSample->normal = normalize(applyNormalMatrix(s, pick.instSlot, lightSample->normal));
    lightSample->emittance = f3(h3.lo.z, h3.lo.w, h3.hi.x);
}

// sampleLight for the RIS candidate loop, fetching the 128-byte light record in steps and stopping as soon as the candidate is
// certain to contribute RGB(0) to this shading point.  ncu: the candidate kernel is bound by the L1 data pipe's wavefront rate
// (every lane reads a different light, so every load instruction of the sampling chain is up to 32 wavefronts) and by issue
// slots; ~70 % of the candidates are dark (65 % lie below the shading horizon).
//   step 0: the bounding sphere of the light triangle - it arrived with the pick (guide entry).  Wholly below the shading
//           horizon -> dark, no further load;
//   step 1: the three vertices (2 x 256 bits) -> sample position.  Below the shading horizon (the BRDFs return RGB(0) when
//           vGiven.z * vSampled.z <= 0)?  -> dark;
//   step 2: the vertex normals, the emittance (256 bits) and the instance's normal matrix (256 + 32 bits) -> un-normalised light
//           normal.  Facing away (lpCos <= 0)? -> dark; otherwise the normalisation: the full sample.
// The tests run on un-normalised vectors with a relative margin of 1e-3 on the cosine (the sphere test with twice that, so
// that every sample point of the triangle would pass the per-sample test too; tests/cpp/sphere_cull_check.cpp) - three orders
// of magnitude above the rounding of the exact evaluation - so borderline samples go on to the exact path and no decision ever
// differs from performDirectLighting's; NaNs and a zero distance fail the comparisons and also go on.  Returns true for "dark"
// (then *areaPDensity is positive); otherwise the outputs are sampleLight's, bit for bit.
GFX_D bool sampleLightUnlessDark(const DevScene &s, float ul, float u0, float u1, const f3 &shadingPoint,
                                 const f3 &shadingNormal, float vOutLocalZ, LightSample* lightSample, float* areaPDensity) {
    const LightPick pick = pickLight(s, ul);
    if (pick.key & kPickNone) {
        *areaPDensity = 0.0f;
        return false;
    }
    {   // step 0: the whole triangle lies below the shading horizon if its bounding sphere does:
        //   -dot(q - p, n) sgn >= t - r |n|  and  |q - p| <= |c - p| + r   for every q in the sphere
        // (radius >= 0 also says that the selection density of this light is a positive finite number)
        const f3 dc = f3(pick.sphere.x, pick.sphere.y, pick.sphere.z) - shadingPoint;
        const float r = pick.sphere.w;
        const float t = -(dot(dc, shadingNormal) * vOutLocalZ);          // > 0: centre on the far side of the surface
        const float margin = t - 1.001f * r * fabsf(vOutLocalZ);
#ifdef GFX_AB_NO_SPHERE
        const bool sphereCull = false;
#else
        const bool sphereCull = true;
#endif
        if (sphereCull && r >= 0.0f && margin > 0.0f && margin * margin > 8e-6f * (sqLength(dc) + r * r) * (vOutLocalZ * vOutLocalZ)) {
            *areaPDensity = 1.0f; // any positive number: the caller only asks whether the density is positive
            return true;
        }
    }
    const float4* e = s.lightTris + kLightTriStride * (size_t)pick.key;
    const F8 h1 = ldg256(e + 2), h2 = ldg256(e + 4);
    const f3 pA(h1.lo.x, h1.lo.y, h1.lo.z), pB(h1.lo.w, h1.hi.x, h1.hi.y), pC(h1.hi.z, h1.hi.w, h2.lo.x);
    float bcA, bcB, bcC;
    squareToTriangle(u0, u1, &bcA, &bcB, &bcC);
    const float density = pick.density;
    *areaPDensity = density;
    const f3 position = bcA * pA + bcB * pB + bcC * pC;

    const float k = 1e-6f; // (1e-3)^2
    const f3 d = position - shadingPoint;
    const float dd = sqLength(d);
    const float b = dot(d, shadingNormal) * vOutLocalZ; // < 0: light and viewer on opposite sides of the surface
    if (density > 0.0f && b < 0.0f && b * b > k * dd * (vOutLocalZ * vOutLocalZ))
        return true;

    const F8 h3 = ldg256(e + 6);
    const f3 nA(h2.lo.y, h2.lo.z, h2.lo.w), nB(h2.hi.x, h2.hi.y, h2.hi.z), nC(h2.hi.w, h3.lo.x, h3.lo.y);
    f3 normal = bcA * nA + bcB * nB + bcC * nC;
    normal = applyNormalMatrix(s, pick.instSlot, normal);
    const float a = dot(d, normal); // > 0: the emitter faces away from the shading point
    if (density > 0.0f && a > 0.0f && a * a > k * dd * sqLength(normal))
        return true;

    lightSample->position = position;
    lightSample->atInfinity = 0;
    lightSample->normal = normalize(normal);
    lightSample->emittance = f3(h3.lo.z, h3.lo.w, h3.hi.x);
    return false;
}

// sampleLightUnlessDark in two halves for the two-phase candidate loop of the ReSTIR initial pass (restir.cu):
//   classifyLight    the pick and step 0 (bounding sphere) -> kLightDark, or the light's key with bit 31 = "the dark tests may
//                    be used" (the selection density is a positive finite number; otherwise the exact path decides);
//   finishLightSample steps 1 and 2 from the key alone (density and instance slot come from the record's H3 copy).
constexpr uint32_t kLightDark = 0xFFFFFFFFu;
constexpr uint32_t kLightCullable = 0x80000000u;
GFX_D uint32_t classifyLight(const DevScene &s, float ul, const F8 &guideEntry, const f3 &shadingPoint, const f3 &shadingNormal,
                             float vOutLocalZ) {
    const LightPick pick = pickLightFromGuide(s, ul, guideEntry);
    if (pick.key & kPickNone)
        return kPickNone; // every "no light" key is the same to the caller
    const f3 dc = f3(pick.sphere.x, pick.sphere.y, pick.sphere.z) - shadingPoint;
    const float r = pick.sphere.w;
    const float t = -(dot(dc, shadingNormal) * vOutLocalZ);
    const float margin = t - 1.001f * r * fabsf(vOutLocalZ);
    if (r >= 0.0f && margin > 0.0f && margin * margin > 8e-6f * (sqLength(dc) + r * r) * (vOutLocalZ * vOutLocalZ))
        return kLightDark;
    return pick.key | (r >= 0.0f ? kLightCullable : 0u);
}
// returns true for "dark"; otherwise *lightSample / *areaPDensity are sampleLight's, bit for bit
GFX_D bool finishLightSample(const DevScene &s, uint32_t code, float u0, float u1, const f3 &shadingPoint, const f3 &shadingNormal,
                             float vOutLocalZ, LightSample* lightSample, float* areaPDensity) {
    if (code & kPickNone) {
        *areaPDensity = 0.0f;
        return false;
    }
    const bool cullable = (code & kLightCullable) != 0; // implies density > 0
    const float4* e = s.lightTris + kLightTriStride * (size_t)(code & 0x3FFFFFFFu);
    const F8 h1 = ldg256(e + 2), h2 = ldg256(e + 4);
    const f3 pA(h1.lo.x, h1.lo.y, h1.lo.z), pB(h1.lo.w, h1.hi.x, h1.hi.y), pC(h1.hi.z, h1.hi.w, h2.lo.x);
    float bcA, bcB, bcC;
    squareToTriangle(u0, u1, &bcA, &bcB, &bcC);
    const f3 position = bcA * pA + bcB * pB + bcC * pC;

    const float k = 1e-6f; // (1e-3)^2
    const f3 d = position - shadingPoint;
    const float dd = sqLength(d);
    const float b = dot(d, shadingNormal) * vOutLocalZ; // < 0: light and viewer on opposite sides of the surface
    if (cullable && b < 0.0f && b * b > k * dd * (vOutLocalZ * vOutLocalZ)) {
        *areaPDensity = 1.0f;
        return true;
    }
    const F8 h3 = ldg256(e + 6);
    const float density = h3.hi.z;
    *areaPDensity = density;
    const f3 nA(h2.lo.y, h2.lo.z, h2.lo.w), nB(h2.hi.x, h2.hi.y, h2.hi.z), nC(h2.hi.w, h3.lo.x, h3.lo.y);
    f3 normal = bcA * nA + bcB * nB + bcC * nC;
    normal = applyNormalMatrix(s, __float_as_uint(h3.hi.w), normal);
    const float a = dot(d, normal); // > 0: the emitter faces away from the shading point
    if (cullable && a > 0.0f && a * a > k * dd * sqLength(normal))
        return true;

    lightSample->position = position;
    lightSample->atInfinity = 0;
    lightSample->normal = normalize(normal);
    lightSample->emittance = f3(h3.lo.z, h3.lo.w, h3.hi.x);
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
