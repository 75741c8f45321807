// wavefront.cuh — the persistent-thread ray tracing kernel shared by trace.cu (C-ABI ray batches, ReSTIR
// visibility queue) and pathtrace.cu (extension + shadow ray queues of the wavefront path tracer).
//
// Why persistent threads: incoherent rays of one warp visit very different numbers of nodes (the shadow
// rays of an 8x4 pixel tile go to 32 different lights), so a ray-per-thread launch idles most lanes while
// the longest traversal finishes.  Here a warp keeps 32 traversal state machines (traverse.cuh) and refills
// finished lanes from a global counter as soon as a quarter of them are idle; one atomic per refill.
//
// Writer concept: `template <bool ANY_HIT, bool STATS> void write(uint32_t ray, const TraversalState&) const`
// is called exactly once per ray, by the lane that traced it.
#pragma once
#include "traverse.cuh"

namespace gfx {

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

static inline int wavefrontGrid() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 8; // 8 x 128 threads = 32 warps per SM
}

} // namespace gfx
