// trace.cu — wavefront ray tracing: persistent threads pulling rays from SoA queues.
//
// Entry points
//  * traceRays            bvh::traverse<8> (common/bvh_builder.cpp:1653-1663) for a batch of 32-byte ray
//                         records -> 32-byte shared::HitObject records (common/common_shared.h:1065-1078);
//  * traceVisibilityQueue the visibility rays of the ReSTIR passes (restir_di_shared.h:559-582 + the AH program
//                         optix_restir_di_kernels.cu:5-8): rays were appended to the frame's queue by the
//                         producing kernel with warp-ballot compaction, results land in a per-pixel byte.
//
// Why persistent threads: incoherent rays of one warp visit very different numbers of nodes (the shadow
// rays of an 8x4 pixel tile go to 32 different lights), so a ray-per-thread launch idles most lanes while
// the longest traversal finishes.  Here a warp keeps 32 traversal state machines (traverse.cuh) and refills
// finished lanes from a global counter as soon as a quarter of them are idle; one atomic per refill.
#include "traverse.cuh"
#include "context.h"

namespace gfx {

struct HitWriter { // GfxHitObject records, indexed by ray
    const uint2* geomToInstMesh;
    uint4* hits;
    template <bool ANY_HIT, bool STATS>
    GFX_D void write(uint32_t ray, const TraversalState &st) const {
        const Hit &h = st.best;
        const bool isHit = h.storageIndex != 0xFFFFFFFFu;
        // STATS: instUserData = internal nodes visited | triangles tested << 16 (bvh::TraversalStatistics)
        const uint32_t userData = STATS ? (min(h.statNodes, 0xFFFFu) | (min(h.statTris, 0xFFFFu) << 16)) : 0u;
        if (ANY_HIT) {
            // visibility payload: dist = 0 when occluded, tmax otherwise (the AH program writes 0.0f)
            hits[2 * (size_t)ray] = make_uint4(__float_as_uint(isHit ? 0.0f : st.tmax), 0xFFFFFFFFu, userData, 0xFFFFFFFFu);
            hits[2 * (size_t)ray + 1] = make_uint4(isHit ? 0u : 0xFFFFFFFFu, 0x7FC00000u, 0x7FC00000u, 0x7FC00000u);
            return;
        }
        const uint32_t inst = isHit && geomToInstMesh ? __ldg(geomToInstMesh + h.geomIndex).x : 0xFFFFFFFFu;
        const float bcA = 1.0f - (h.bcB + h.bcC);
        hits[2 * (size_t)ray] = make_uint4(__float_as_uint(h.dist), inst, userData, h.geomIndex);
        hits[2 * (size_t)ray + 1] = isHit
            ? make_uint4(h.primIndex, __float_as_uint(bcA), __float_as_uint(h.bcB), __float_as_uint(h.bcC))
            : make_uint4(0xFFFFFFFFu, 0x7FC00000u, 0x7FC00000u, 0x7FC00000u);
    }
};

struct VisibilityWriter { // one byte per pixel that asked
    const uint32_t* rayPixel;
    uint8_t* visibility;
    template <bool ANY_HIT, bool STATS>
    GFX_D void write(uint32_t ray, const TraversalState &st) const {
        visibility[__ldg(rayPixel + ray)] = st.best.storageIndex == 0xFFFFFFFFu ? 1 : 0;
    }
};

template <bool ANY_HIT, bool STATS, typename Writer>
__global__ void __launch_bounds__(128) k_traceWavefront(DevBvh bvh, const float4* __restrict__ rays,
                                                        const uint32_t* __restrict__ numRaysPtr, uint32_t numRaysImm,
                                                        uint32_t* __restrict__ fetchCounter, Writer writer) {
    const uint32_t total = numRaysPtr ? *numRaysPtr : numRaysImm;
    if (total == 0 || bvh.numNodes == 0) {
        // nothing to traverse: every ray misses
        if (bvh.numNodes == 0) {
            TraversalState st;
            for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
                const float4 r1 = __ldg(rays + 2 * (size_t)r + 1);
                traverseInit(st, f3(0, 0, 0), f3(0, 0, 1), 0.0f, r1.w);
                writer.template write<ANY_HIT, STATS>(r, st);
            }
        }
        return;
    }
    const uint32_t lane = threadIdx.x & 31u;
    TraversalState st;
    bool active = false;
    uint32_t myRay = 0;
    while (true) {
        // ---- refill idle lanes (one atomic per warp)
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader)
                base = atomicAdd(fetchCounter, (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active) {
                const uint32_t r = base + __popc(idle & ((1u << lane) - 1u));
                if (r < total) {
                    const float4 r0 = __ldg(rays + 2 * (size_t)r);
                    const float4 r1 = __ldg(rays + 2 * (size_t)r + 1);
                    traverseInit(st, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), r0.w, r1.w);
                    myRay = r;
                    active = true;
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, active))
            break;
        // ---- advance every active lane by up to 8 nodes, leave early once a quarter of the warp is idle
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            if (active && !traverseStep<ANY_HIT, STATS>(bvh, st)) {
                writer.template write<ANY_HIT, STATS>(myRay, st);
                active = false;
            }
            if (__popc(__ballot_sync(0xFFFFFFFFu, active)) <= 24)
                break;
        }
    }
}

static int wavefrontGrid() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 8; // 8 x 128 threads = 32 warps per SM
}

int traceRays(gfx_ctx* ctx, cudaStream_t stream, const GfxRay* dRays, uint32_t numRays, GfxHitObject* dHits, int mode) {
    if (numRays == 0)
        return GFX_OK;
    const DevScene dev = ctx->devScene();
    const float4* r = reinterpret_cast<const float4*>(dRays);
    HitWriter w{ dev.geomToInstMesh, reinterpret_cast<uint4*>(dHits) };
    uint32_t* counter = ctx->traceFetchCounter;
    GFX_CUDA(ctx, cudaMemsetAsync(counter, 0, 4, stream));
    const int grid = min(wavefrontGrid(), (int)((numRays + 127) / 128));
    switch (mode) {
    case GFX_TRACE_CLOSEST: k_traceWavefront<false, false><<<grid, 128, 0, stream>>>(dev.bvh, r, nullptr, numRays, counter, w); break;
    case GFX_TRACE_ANY: k_traceWavefront<true, false><<<grid, 128, 0, stream>>>(dev.bvh, r, nullptr, numRays, counter, w); break;
    case GFX_TRACE_CLOSEST | GFX_TRACE_STATS: k_traceWavefront<false, true><<<grid, 128, 0, stream>>>(dev.bvh, r, nullptr, numRays, counter, w); break;
    case GFX_TRACE_ANY | GFX_TRACE_STATS: k_traceWavefront<true, true><<<grid, 128, 0, stream>>>(dev.bvh, r, nullptr, numRays, counter, w); break;
    default:
        ctx->setError("gfx_trace: unknown mode");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

int resetVisibilityQueue(gfx_ctx* ctx, cudaStream_t stream) {
    GFX_CUDA(ctx, cudaMemsetAsync(ctx->frame.rayCounters, 0, 8, stream));
    return GFX_OK;
}

int traceVisibilityQueue(gfx_ctx* ctx, cudaStream_t stream) {
    const DevScene dev = ctx->devScene();
    const FrameState &F = ctx->frame;
    VisibilityWriter w{ F.rayPixel, F.visibility };
    k_traceWavefront<true, false><<<wavefrontGrid(), 128, 0, stream>>>(dev.bvh, F.rayQueue, F.rayCounters, 0u, F.rayCounters + 1, w);
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
