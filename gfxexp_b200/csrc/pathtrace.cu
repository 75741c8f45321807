// pathtrace.cu — the unidirectional path tracer as a wavefront pipeline on sm_100a.
//
// Replaces pathTracing.optixPipeline.launch(W, H, 1) (path_tracing/path_tracing_main.cpp:1780-1789) and its
// programs pathTrace_rayGen_generic / pathTrace_closestHit_generic / miss / performNextEventEstimation
// (path_tracing/gpu_kernels/optix_pathtracing_kernels.cu:18-71, 73-216, 218-300, 310-341) with
// computeSurfacePoint<> (path_tracing/path_tracing_shared.h:484-621); compile-time switches as in :12-16
// (area sampling, implicit + explicit light sampling, MIS).
//
// OptiX runs this as a per-pixel megakernel that recurses into closest-hit programs.  Here a path is a
// record in HBM and one frame is
//     k_ptFirstHit                          shade the G-buffer hit: emission, NEE (shadow ray request), BSDF sample
//     for pathLength = 2 .. maxPathLength
//         k_traceWavefront<any>             shadow rays of the previous vertex; the writer adds the unoccluded
//                                           NEE contribution to the path's radiance (keeps the reference's
//                                           summation order: NEE of vertex k before the emission of vertex k+1)
//         k_traceWavefront<closest>         extension rays -> (geometry, primitive, barycentrics) per queue slot
//         k_ptBounce                        closest-hit program on the compacted queue of live paths
//     k_ptAccumulate                        running mean into the beauty buffer
// Queues are compacted with one atomic per warp (ballot), each round has its own counters (one memset per
// frame) and the trace kernel is the persistent-thread kernel of wavefront.cuh, so lanes that finish early or
// whose path died do not idle.  Per-path state: rng (frame.rng), alpha|prevDirPDensity and radiance as two
// float4 planes indexed by pixel.  The RNG draw order of every pixel is the reference's.
#include "wavefront.cuh"
#include "lighting.cuh"
#include "context.h"

namespace gfx {

constexpr uint32_t kMaxPathRounds = 32;

struct DevPathState {
    float4* alphaPdf;      // per pixel: alpha.rgb, prevDirPDensity
    float4* radiance;      // per pixel: contribution.rgb
    float4* extRays[2];    // per slot: (org, tmin) (dir, tmax)
    uint32_t* extPixel[2]; // per slot: pixel index
    uint4* extHits;        // per slot: geomIndex, primIndex, bcB, bcC
    float4* shadowRays;    // per slot
    uint32_t* shadowPixel; // per slot
    float4* shadowPending; // per slot: alpha * NEE value if unoccluded, w = scale (misWeight / areaPDensity)
    uint32_t* counters;    // per round: extCount, extFetch, shadowCount, shadowFetch
};

struct SurfacePoint {
    f3 positionInWorld, shadingNormalInWorld, texCoord0DirInWorld, geometricNormalInWorld;
    float hypAreaPDensity;
};

struct TriangleVertices {
    f3 pA, pB, pC, nA, nB, nC, tA, tB, tC;
};
GFX_D TriangleVertices fetchTriangle(const DevScene &s, const DevMesh &mesh, uint32_t primIndex) {
    const uint4 tri = __ldg(s.triangles + mesh.triBase + primIndex);
    const float4* vA = s.vertices + 3 * (size_t)(mesh.vertexBase + tri.x);
    const float4* vB = s.vertices + 3 * (size_t)(mesh.vertexBase + tri.y);
    const float4* vC = s.vertices + 3 * (size_t)(mesh.vertexBase + tri.z);
    const float4 a0 = __ldg(vA), a1 = __ldg(vA + 1), a2 = __ldg(vA + 2);
    const float4 b0 = __ldg(vB), b1 = __ldg(vB + 1), b2 = __ldg(vB + 2);
    const float4 c0 = __ldg(vC), c1 = __ldg(vC + 1), c2 = __ldg(vC + 2);
    TriangleVertices t;
    t.pA = f3(a0.x, a0.y, a0.z); t.pB = f3(b0.x, b0.y, b0.z); t.pC = f3(c0.x, c0.y, c0.z);
    t.nA = f3(a1.x, a1.y, a1.z); t.nB = f3(b1.x, b1.y, b1.z); t.nC = f3(c1.x, c1.y, c1.z);
    t.tA = f3(a2.x, a2.y, a2.z); t.tB = f3(b2.x, b2.y, b2.z); t.tC = f3(c2.x, c2.y, c2.z);
    return t;
}

// path_tracing_shared.h:582-621 (first hit, from GBuffer0's quantised barycentrics)
GFX_D void computeSurfacePointFromGBuffer(const DevScene &s, const DevInstance* inst, const DevMesh &mesh,
                                          uint32_t primIndex, float bcB, float bcC, SurfacePoint* sp) {
    const TriangleVertices t = fetchTriangle(s, mesh, primIndex);
    const float bcA = 1 - (bcB + bcC);
    const f3 positionInObj = bcA * t.pA + bcB * t.pB + bcC * t.pC;
    sp->positionInWorld = xfmPoint(inst->transform, positionInObj);
    sp->geometricNormalInWorld = normalize(mul3x3(inst->normalMatrix, cross(t.pB - t.pA, t.pC - t.pA)));
    const f3 shadingNormalInObj = bcA * t.nA + bcB * t.nB + bcC * t.nC;
    const f3 texCoord0DirInObj = bcA * t.tA + bcB * t.tB + bcC * t.tC;
    sp->shadingNormalInWorld = normalize(mul3x3(inst->normalMatrix, shadingNormalInObj));
    sp->texCoord0DirInWorld = xfmVector(inst->transform, texCoord0DirInObj);
    sp->texCoord0DirInWorld = normalize(
        sp->texCoord0DirInWorld - dot(sp->shadingNormalInWorld, sp->texCoord0DirInWorld) * sp->shadingNormalInWorld);
    if (!allFinite(sp->shadingNormalInWorld)) {
        sp->geometricNormalInWorld = f3(0, 0, 1);
        sp->shadingNormalInWorld = f3(0, 0, 1);
        sp->texCoord0DirInWorld = f3(1, 0, 0);
    }
    if (!allFinite(sp->texCoord0DirInWorld)) {
        f3 bitangent;
        makeCoordinateSystem(sp->shadingNormalInWorld, &sp->texCoord0DirInWorld, &bitangent);
    }
    sp->hypAreaPDensity = 0.0f;
}

// path_tracing_shared.h:484-580 with computeHypotheticalAreaPDensity = true, useSolidAngleSampling = false
GFX_D void computeSurfacePointAtHit(const DevScene &s, const DevInstance* inst, const DevMesh &mesh,
                                    uint32_t primIndex, float bcB, float bcC, SurfacePoint* sp) {
    const TriangleVertices t = fetchTriangle(s, mesh, primIndex);
    const f3 pA = xfmPoint(inst->transform, t.pA);
    const f3 pB = xfmPoint(inst->transform, t.pB);
    const f3 pC = xfmPoint(inst->transform, t.pC);
    const float bcA = 1 - (bcB + bcC);

    sp->positionInWorld = bcA * pA + bcB * pB + bcC * pC;
    const f3 shadingNormalInObj = bcA * t.nA + bcB * t.nB + bcC * t.nC;
    const f3 texCoord0DirInObj = bcA * t.tA + bcB * t.tB + bcC * t.tC;

    sp->geometricNormalInWorld = cross(pB - pA, pC - pA);
    const float area = 0.5f * length(sp->geometricNormalInWorld);
    sp->geometricNormalInWorld = sp->geometricNormalInWorld / (2 * area);

    sp->shadingNormalInWorld = normalize(mul3x3(inst->normalMatrix, shadingNormalInObj));
    sp->texCoord0DirInWorld = normalize(xfmVector(inst->transform, texCoord0DirInObj));
    if (!allFinite(sp->shadingNormalInWorld)) {
        sp->shadingNormalInWorld = f3(0, 0, 1);
        sp->texCoord0DirInWorld = f3(1, 0, 0);
    }
    if (!allFinite(sp->texCoord0DirInWorld)) {
        f3 bitangent;
        makeCoordinateSystem(sp->shadingNormalInWorld, &sp->texCoord0DirInWorld, &bitangent);
    }

    // hypothetical density with which explicit light sampling would have produced this point (:535-575)
    float lightProb = 1.0f;
    const float instImportance = inst->geomIntegral;
    lightProb *= (pow2f(inst->uniformScale) * instImportance) / __ldg(s.instIntegral);
    lightProb *= mesh.primIntegral / instImportance;
    if (!isfinite(lightProb)) {
        sp->hypAreaPDensity = 0.0f;
        return;
    }
    lightProb *= mesh.primIntegral == 0.0f ? 0.0f : __ldg(s.primWeights + mesh.triBase + primIndex) / mesh.primIntegral;
    sp->hypAreaPDensity = lightProb / area;
}

// one slot of a compacted queue per requesting lane; one atomic per warp
GFX_D uint32_t allocQueueSlot(uint32_t* counter, bool want, uint32_t lane) {
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, want);
    if (mask == 0)
        return 0;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader)
        base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// The part of a path vertex both programs share: next event estimation (optix_pathtracing_kernels.cu:18-71)
// and BSDF sampling of the next direction (:136-146 / :283-299).  Rays are only requested here.
struct VertexOutput {
    bool wantShadow, wantExtension;
    f3 shadowDir;
    float shadowTmax;
    float4 pending;
    f3 nextDir;
    f3 alpha;
    float dirPDensity;
};

GFX_D void shadeVertex(const DevScene &s, const f3 &positionInWorld, const f3 &vOutLocal, const ReferenceFrame &shadingFrame,
                       const BSDF &bsdf, PCG32RNG &rng, f3 alpha, f3* radiance, VertexOutput* out) {
    // ---- next event estimation
    out->wantShadow = false;
    {
        const float uLight = rng.getFloat0cTo1o();
        const float u0 = rng.getFloat0cTo1o();
        const float u1 = rng.getFloat0cTo1o();
        LightSample lightSample;
        float areaPDensity = 0.0f;
        sampleLight(s, uLight, u0, u1, &lightSample, &areaPDensity);
        if (areaPDensity > 0.0f) {
            f3 shadowRay = lightSample.position - positionInWorld;
            const float dist2 = sqLength(shadowRay);
            const float dist = sqrtf(dist2);
            shadowRay /= dist;
            const f3 vInLocal = shadingFrame.toLocal(shadowRay);
            const float lpCos = fabsf(dot(shadowRay, lightSample.normal));
            float bsdfPDensity = bsdf.evaluatePDF(vOutLocal, vInLocal) * lpCos / dist2;
            if (!isfinite(bsdfPDensity))
                bsdfPDensity = 0.0f;
            const float lightPDensity = areaPDensity;
            const float misWeight = pow2f(lightPDensity) / (pow2f(bsdfPDensity) + pow2f(lightPDensity));
            const float scale = misWeight / areaPDensity;
            // performDirectLighting<.., true> = visibility * value; the ray is only needed when value != 0
            const f3 value = performDirectLighting<false>(s, positionInWorld, vOutLocal, shadingFrame, bsdf, lightSample);
            const f3 unoccluded = alpha * (value * scale);
            if (value.x == 0.0f && value.y == 0.0f && value.z == 0.0f) {
                *radiance += unoccluded; // 0 (or NaN for a non-finite scale) whatever the visibility
            }
            else {
                out->wantShadow = true;
                out->shadowDir = shadowRay;
                out->shadowTmax = dist * 0.9999f;
                out->pending = make_float4(unoccluded.x, unoccluded.y, unoccluded.z, scale);
            }
        }
    }
    // ---- next direction
    f3 vInLocal;
    float dirPDensity;
    const float uDir0 = rng.getFloat0cTo1o();
    const float uDir1 = rng.getFloat0cTo1o();
    alpha *= bsdf.sampleThroughput(vOutLocal, uDir0, uDir1, &vInLocal, &dirPDensity);
    out->nextDir = shadingFrame.fromLocal(vInLocal);
    out->alpha = alpha;
    out->dirPDensity = dirPDensity;
    // the path extension loop (:161-165) stops on an invalid sample before tracing anything
    out->wantExtension = dirPDensity > 0.0f && isfinite(dirPDensity);
}

GFX_D void emitRays(const DevScene &s, const DevPathState &ps, uint32_t* roundCounters, uint32_t nextQueue, uint32_t lane,
                    uint32_t pix, const f3 &positionInWorld, const VertexOutput &v, bool alive) {
    const bool wantShadow = alive && v.wantShadow;
    const bool wantExt = alive && v.wantExtension;
    const uint32_t shadowSlot = allocQueueSlot(roundCounters + 2, wantShadow, lane);
    const uint32_t extSlot = allocQueueSlot(roundCounters + 0, wantExt, lane);
    if (wantShadow) {
        ps.shadowRays[2 * (size_t)shadowSlot] = make_float4(positionInWorld.x, positionInWorld.y, positionInWorld.z, 0.0f);
        ps.shadowRays[2 * (size_t)shadowSlot + 1] = make_float4(v.shadowDir.x, v.shadowDir.y, v.shadowDir.z, v.shadowTmax);
        ps.shadowPixel[shadowSlot] = pix;
        ps.shadowPending[shadowSlot] = v.pending;
    }
    if (wantExt) {
        ps.extRays[nextQueue][2 * (size_t)extSlot] = make_float4(positionInWorld.x, positionInWorld.y, positionInWorld.z, 0.0f);
        ps.extRays[nextQueue][2 * (size_t)extSlot + 1] = make_float4(v.nextDir.x, v.nextDir.y, v.nextDir.z, 3.402823466e+38f);
        ps.extPixel[nextQueue][extSlot] = pix;
    }
    // ray statistics (frame.stats[0]): one atomic per warp
    const uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, wantShadow)) + __popc(__ballot_sync(0xFFFFFFFFu, wantExt));
    if (lane == 0 && n)
        atomicAdd(s.rayCounter, (unsigned long long)n);
}

// pathTrace_rayGen_generic up to the path extension loop (:73-160)
__global__ void __launch_bounds__(64) k_ptFirstHit(DevScene s, DevFrame f, DevFrameParams p, DevPathState ps) {
    const uint32_t x = blockIdx.x * 8 + threadIdx.x;
    const uint32_t y = p.y0 + blockIdx.y * 8 + threadIdx.y;
    const uint32_t lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31u;
    const bool inside = x < f.W && y < p.y1;
    const uint32_t pix = inside ? y * f.W + x : 0u;

    bool alive = false;
    f3 positionInWorld(0.0f);
    VertexOutput v;
    v.wantShadow = v.wantExtension = false;
    if (inside) {
        const uint4 gb0 = f.gb0[p.bufferIndex][pix];
        f3 radiance(0.001f, 0.001f, 0.001f);
        if (gb0.x != 0xFFFFFFFFu) {
            const DevInstance* inst = s.instances + gb0.x;
            const DevMesh mesh = s.meshes[gb0.y];
            const float bcB = decodeBarycentric((uint16_t)(gb0.w & 0xFFFFu));
            const float bcC = decodeBarycentric((uint16_t)(gb0.w >> 16));
            SurfacePoint sp;
            computeSurfacePointFromGBuffer(s, inst, mesh, gb0.z, bcB, bcC, &sp);

            const f3 alpha(1.0f);
            PCG32RNG rng{ f.rng[pix] };
            const GfxMaterialDesc* mat = s.materials + mesh.materialSlot;
            const f3 vOut = normalize(p.camera.position - sp.positionInWorld);
            const float frontHit = dot(vOut, sp.geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
            positionInWorld = offsetRayOrigin(sp.positionInWorld, frontHit * sp.geometricNormalInWorld);
            const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
            const f3 vOutLocal = shadingFrame.toLocal(vOut);

            radiance = f3(0.0f);
            if (vOutLocal.z > 0 && mat->hasEmittance)
                radiance += alpha * f3(mat->emittance[0], mat->emittance[1], mat->emittance[2]) / kPi;
            const BSDF bsdf = setupBsdf(s, mesh.materialSlot);
            shadeVertex(s, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, alpha, &radiance, &v);
            f.rng[pix] = rng.state;
            ps.alphaPdf[pix] = make_float4(v.alpha.x, v.alpha.y, v.alpha.z, v.dirPDensity);
            alive = true;
        }
        ps.radiance[pix] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
    }
    emitRays(s, ps, ps.counters, 0u, lane, pix, positionInWorld, v, alive);
}

// pathTrace_closestHit_generic (:218-300) + the bookkeeping of the path extension loop (:161-194) for the
// compacted queue of live paths; `round` = pathLength - 2 selects the queue parity and the counters.
__global__ void __launch_bounds__(64) k_ptBounce(DevScene s, DevFrame f, DevPathState ps, uint32_t round, uint32_t maxLengthTerminate) {
    const uint32_t curQueue = round & 1u, nextQueue = curQueue ^ 1u;
    const uint32_t* roundCounters = ps.counters + 4 * round;
    uint32_t* nextCounters = ps.counters + 4 * (round + 1);
    const uint32_t count = roundCounters[0];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const float initImportance = sRGB_calcLuminance(f3(1.0f));

    for (uint32_t base = warp * 32u; base < count; base += warpsPerGrid * 32u) {
        const uint32_t slot = base + lane;
        bool alive = false;
        uint32_t pix = 0;
        f3 positionInWorld(0.0f);
        VertexOutput v;
        v.wantShadow = v.wantExtension = false;
        if (slot < count) {
            const uint4 hit = ps.extHits[slot];
            if (hit.y != 0xFFFFFFFFu) { // miss program: no environment light, the path ends
                pix = ps.extPixel[curQueue][slot];
                const float4 r0 = ps.extRays[curQueue][2 * (size_t)slot];
                const float4 r1 = ps.extRays[curQueue][2 * (size_t)slot + 1];
                const f3 rayOrigin(r0.x, r0.y, r0.z), rayDir(r1.x, r1.y, r1.z);
                const float4 ap = ps.alphaPdf[pix];
                f3 alpha(ap.x, ap.y, ap.z);
                const float prevDirPDensity = ap.w;
                const float4 rad = ps.radiance[pix];
                f3 radiance(rad.x, rad.y, rad.z);
                PCG32RNG rng{ f.rng[pix] };

                const uint2 im = __ldg(s.geomToInstMesh + hit.x);
                const DevInstance* inst = s.instances + im.x;
                const DevMesh mesh = s.meshes[im.y];
                SurfacePoint sp;
                computeSurfacePointAtHit(s, inst, mesh, hit.y, __uint_as_float(hit.z), __uint_as_float(hit.w), &sp);
                const GfxMaterialDesc* mat = s.materials + mesh.materialSlot;

                const f3 vOut = normalize(-rayDir);
                const float frontHit = dot(vOut, sp.geometricNormalInWorld) >= 0.0f ? 1.0f : -1.0f;
                const ReferenceFrame shadingFrame(sp.shadingNormalInWorld, sp.texCoord0DirInWorld);
                positionInWorld = offsetRayOrigin(sp.positionInWorld, frontHit * sp.geometricNormalInWorld);
                const f3 vOutLocal = shadingFrame.toLocal(vOut);

                // implicit light sampling, MIS against the hypothetical NEE density (:262-275)
                if (vOutLocal.z > 0 && mat->hasEmittance) {
                    const f3 emittance(mat->emittance[0], mat->emittance[1], mat->emittance[2]);
                    const float dist2 = sqLength(positionInWorld - rayOrigin);
                    const float lightPDensity = sp.hypAreaPDensity * dist2 / vOutLocal.z;
                    const float bsdfPDensity = prevDirPDensity;
                    const float misWeight = pow2f(bsdfPDensity) / (pow2f(bsdfPDensity) + pow2f(lightPDensity));
                    radiance += alpha * emittance * (misWeight / kPi);
                }

                // Russian roulette (:277-281)
                const float continueProb = fminf(sRGB_calcLuminance(alpha) / initImportance, 1.0f);
                if (!(rng.getFloat0cTo1o() >= continueProb || maxLengthTerminate)) {
                    alpha /= continueProb;
                    const BSDF bsdf = setupBsdf(s, mesh.materialSlot);
                    shadeVertex(s, positionInWorld, vOutLocal, shadingFrame, bsdf, rng, alpha, &radiance, &v);
                    ps.alphaPdf[pix] = make_float4(v.alpha.x, v.alpha.y, v.alpha.z, v.dirPDensity);
                    alive = true;
                }
                f.rng[pix] = rng.state;
                ps.radiance[pix] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
            }
        }
        emitRays(s, ps, nextCounters, nextQueue, lane, pix, positionInWorld, v, alive);
    }
}

// ray-gen epilogue (:202-215)
__global__ void __launch_bounds__(256) k_ptAccumulate(DevFrame f, DevFrameParams p, DevPathState ps) {
    const uint32_t n = (p.y1 - p.y0) * f.W;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const size_t pix = (size_t)p.y0 * f.W + i;
    const float4 rad = ps.radiance[pix];
    const f3 contribution(rad.x, rad.y, rad.z);
    f3 prevColorResult(0.0f);
    if (p.numAccumFrames > 0) {
        const float4 b = f.beauty[pix];
        prevColorResult = f3(b.x, b.y, b.z);
    }
    const float curWeight = 1.0f / (1 + p.numAccumFrames);
    const f3 colorResult = (1 - curWeight) * prevColorResult + curWeight * contribution;
    f.beauty[pix] = make_float4(colorResult.x, colorResult.y, colorResult.z, 1.0f);
}

struct ExtensionHitWriter { // closest hit -> what the closest-hit program reads from OptiX (HitPointParameter::get)
    uint4* hits;
    template <bool ANY_HIT, bool STATS>
    GFX_D void write(uint32_t ray, const TraversalState &st) const {
        const Hit &h = st.best;
        const bool isHit = h.storageIndex != 0xFFFFFFFFu;
        hits[ray] = make_uint4(h.geomIndex, isHit ? h.primIndex : 0xFFFFFFFFu, __float_as_uint(h.bcB), __float_as_uint(h.bcC));
    }
};

struct ShadowAccumulateWriter { // visibility * (alpha * f * Le * G * misWeight / p) of performNextEventEstimation
    const uint32_t* shadowPixel;
    const float4* pending;
    float4* radiance;
    template <bool ANY_HIT, bool STATS>
    GFX_D void write(uint32_t ray, const TraversalState &st) const {
        const float4 c = pending[ray];
        const bool unoccluded = st.best.storageIndex == 0xFFFFFFFFu;
        if (!unoccluded && isfinite(c.w))
            return; // alpha * (0 * scale) = 0
        float4* dst = radiance + shadowPixel[ray]; // at most one shadow ray per path and round: no race
        float4 r = *dst;
        if (unoccluded) {
            r.x += c.x; r.y += c.y; r.z += c.z;
        }
        else {
            const float nan = __int_as_float(0x7FC00000);
            r.x += nan; r.y += nan; r.z += nan;
        }
        *dst = r;
    }
};

int launchPathTrace(gfx_ctx* ctx, cudaStream_t stream, const GfxFrameParams* params, int variant) {
    if (variant != GFX_PT_BASELINE) {
        ctx->setError("gfx_pathtrace_launch: unknown variant");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    FrameState &F = ctx->frame;
    const size_t n = (size_t)F.W * F.H;
    if (!F.ptAlphaPdf) {
        GFX_CUDA(ctx, cudaMalloc(&F.ptAlphaPdf, n * 16));
        GFX_CUDA(ctx, cudaMalloc(&F.ptRadiance, n * 16));
        for (int i = 0; i < 2; ++i) {
            GFX_CUDA(ctx, cudaMalloc(&F.ptExtRays[i], n * 32));
            GFX_CUDA(ctx, cudaMalloc(&F.ptExtPixel[i], n * 4));
        }
        GFX_CUDA(ctx, cudaMalloc(&F.ptExtHits, n * 16));
        GFX_CUDA(ctx, cudaMalloc(&F.ptShadowPending, n * 16));
        GFX_CUDA(ctx, cudaMalloc(&F.ptCounters, (kMaxPathRounds + 1) * 16));
    }
    const DevFrameParams p = makeDevParams(ctx, params);
    if (p.y1 <= p.y0)
        return GFX_OK;
    const DevScene s = ctx->devScene();
    const DevFrame f = ctx->devFrame();
    DevPathState ps;
    ps.alphaPdf = F.ptAlphaPdf;
    ps.radiance = F.ptRadiance;
    for (int i = 0; i < 2; ++i) {
        ps.extRays[i] = F.ptExtRays[i];
        ps.extPixel[i] = F.ptExtPixel[i];
    }
    ps.extHits = F.ptExtHits;
    ps.shadowRays = F.rayQueue;
    ps.shadowPixel = F.rayPixel;
    ps.shadowPending = F.ptShadowPending;
    ps.counters = F.ptCounters;

    const uint32_t maxPathLength = params->maxPathLength ? params->maxPathLength : 5u;
    const uint32_t numRounds = min(max(maxPathLength, 2u) - 1u, kMaxPathRounds); // pathLength = 2 .. maxPathLength
    GFX_CUDA(ctx, cudaMemsetAsync(F.ptCounters, 0, (kMaxPathRounds + 1) * 16, stream));

    const dim3 block(8, 8);
    const dim3 grid((F.W + 7) / 8, (p.y1 - p.y0 + 7) / 8);
    k_ptFirstHit<<<grid, block, 0, stream>>>(s, f, p, ps);
    ctx->launches++;
    const int traceGrid = wavefrontGrid();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const ShadowAccumulateWriter shadowWriter{ ps.shadowPixel, ps.shadowPending, ps.radiance };
    const ExtensionHitWriter extWriter{ ps.extHits };
    for (uint32_t round = 0; round < numRounds; ++round) {
        uint32_t* c = F.ptCounters + 4 * round;
        k_traceWavefront<true, false><<<traceGrid, 128, 0, stream>>>(s.bvh, ps.shadowRays, c + 2, 0u, c + 3, shadowWriter);
        k_traceWavefront<false, false><<<traceGrid, 128, 0, stream>>>(s.bvh, ps.extRays[round & 1], c + 0, 0u, c + 1, extWriter);
        const uint32_t maxLengthTerminate = (round + 2 >= maxPathLength || round + 1 == numRounds) ? 1u : 0u;
        k_ptBounce<<<sms * 16, 64, 0, stream>>>(s, f, ps, round, maxLengthTerminate);
        ctx->launches += 3;
    }
    // the last round requests no rays (maxLengthTerminate), nothing is left in flight
    const uint32_t numPixels = (p.y1 - p.y0) * F.W;
    k_ptAccumulate<<<(numPixels + 255) / 256, 256, 0, stream>>>(f, p, ps);
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
