// trace.cu — wavefront ray batch entry point: bvh::traverse<8> (common/bvh_builder.cpp:1653-1663)
// for SoA-free 32-byte ray records in HBM -> 32-byte shared::HitObject records
// (common/common_shared.h:1065-1078).  One thread per ray; rays and hits are read/written as two
// aligned 16-byte vectors, so a warp moves 1 KiB per access.
#include "traverse.cuh"
#include "context.h"

namespace gfx {

template <bool ANY_HIT, bool STATS>
__global__ void __launch_bounds__(128) k_trace(DevBvh bvh, const uint2* __restrict__ geomToInstMesh,
                                               const float4* __restrict__ rays, uint32_t numRays,
                                               uint4* __restrict__ hits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numRays)
        return;
    const float4 r0 = __ldg(rays + 2 * (size_t)i);
    const float4 r1 = __ldg(rays + 2 * (size_t)i + 1);
    const Hit h = traverseBvh<ANY_HIT, STATS>(bvh, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), r0.w, r1.w);
    const bool isHit = h.storageIndex != 0xFFFFFFFFu;
    // STATS: instUserData = internal nodes visited | triangles tested << 16 (bvh::TraversalStatistics)
    const uint32_t userData = STATS ? (min(h.statNodes, 0xFFFFu) | (min(h.statTris, 0xFFFFu) << 16)) : 0u;
    if (ANY_HIT) {
        // visibility payload: dist = 0 when occluded, tmax otherwise (the AH program writes 0.0f,
        // restir_di/gpu_kernels/optix_restir_di_kernels.cu:5-8)
        hits[2 * (size_t)i] = make_uint4(__float_as_uint(isHit ? 0.0f : r1.w), 0xFFFFFFFFu, userData, 0xFFFFFFFFu);
        hits[2 * (size_t)i + 1] = make_uint4(isHit ? 0u : 0xFFFFFFFFu, 0x7FC00000u, 0x7FC00000u, 0x7FC00000u);
        return;
    }
    const uint32_t inst = isHit && geomToInstMesh ? __ldg(geomToInstMesh + h.geomIndex).x : 0xFFFFFFFFu;
    const float bcA = 1.0f - (h.bcB + h.bcC);
    hits[2 * (size_t)i] = make_uint4(__float_as_uint(h.dist), inst, userData, h.geomIndex);
    hits[2 * (size_t)i + 1] = isHit
        ? make_uint4(h.primIndex, __float_as_uint(bcA), __float_as_uint(h.bcB), __float_as_uint(h.bcC))
        : make_uint4(0xFFFFFFFFu, 0x7FC00000u, 0x7FC00000u, 0x7FC00000u);
}

int traceRays(gfx_ctx* ctx, cudaStream_t stream, const GfxRay* dRays, uint32_t numRays, GfxHitObject* dHits, int mode) {
    if (numRays == 0)
        return GFX_OK;
    const DevScene dev = ctx->devScene();
    const uint32_t blocks = (numRays + 127) / 128;
    const float4* r = reinterpret_cast<const float4*>(dRays);
    uint4* h = reinterpret_cast<uint4*>(dHits);
    switch (mode) {
    case GFX_TRACE_CLOSEST: k_trace<false, false><<<blocks, 128, 0, stream>>>(dev.bvh, dev.geomToInstMesh, r, numRays, h); break;
    case GFX_TRACE_ANY: k_trace<true, false><<<blocks, 128, 0, stream>>>(dev.bvh, dev.geomToInstMesh, r, numRays, h); break;
    case GFX_TRACE_CLOSEST | GFX_TRACE_STATS: k_trace<false, true><<<blocks, 128, 0, stream>>>(dev.bvh, dev.geomToInstMesh, r, numRays, h); break;
    case GFX_TRACE_ANY | GFX_TRACE_STATS: k_trace<true, true><<<blocks, 128, 0, stream>>>(dev.bvh, dev.geomToInstMesh, r, numRays, h); break;
    default:
        ctx->setError("gfx_trace: unknown mode");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
