// peer.cu — one-sided seam-row exchange between the strip-sharded GPUs of one box over NVLink peer memory.
//
// Where the reference has nothing (it is a single-GPU program; SURVEY.md §8e shards its frame by screen strips), a rank's
// spatial-reuse pass reads reservoirs up to spatialNeighborRadius rows across the seam (restir_di_main.cpp:2069) and its
// temporal pass reads the previous frame's final reservoirs across it.  Every rank allocates full-frame buffers and
// addresses rows globally, so a seam row has the same offset on both sides: the producer copies its rows straight into
// the neighbour's buffer (cudaIpc-mapped, stores travel over NVLink / NVSwitch) and then raises a sequence flag in the
// neighbour's memory; the consumer's stream waits on its own flag.  No host round trip, no rendezvous, ~4 small launches
// per seam instead of a grouped NCCL send/recv of every plane.
//
// Ordering: the data kernel(s) and the signal kernel are enqueued on one stream; the signal kernel issues a system-scope
// fence before a release store of the sequence number, the wait kernel polls with acquire loads.  A rank can only run one
// exchange ahead of its neighbour (every exchange waits for the neighbour's flag of the same sequence number), and
// consecutive exchanges target different reservoir buffers, so pushed rows are never overwritten while still being read.
#include "context.h"
#include <dlfcn.h>
#include <map>

namespace gfx {

constexpr uint32_t kPeerFlagWords = 64;      // [0] from the upper neighbour, [1] from the lower one, [63] = time-out marker
constexpr int kPeerFlagsBufferId = -1;

struct PeerLink {
    std::map<std::pair<int, uint32_t>, void*> buffers; // (bufferId, index) -> mapped base pointer
    uint32_t* flags = nullptr;
    std::vector<void*> mapped;
};
struct PeerState {
    uint32_t* localFlags = nullptr;
    PeerLink links[2];
};
static std::map<gfx_ctx*, PeerState> g_peers;

__global__ void k_peerPushRows(uint2* __restrict__ dst, const uint2* __restrict__ src, uint32_t planes, size_t planeStride,
                               size_t rowOffset, size_t count) {
    // dst / src are 8-byte element views of the same buffer on two GPUs; copy `count` elements of every plane
    const size_t total = count * planes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t plane = i / count, e = i - plane * count;
        const size_t at = plane * planeStride + rowOffset + e;
        dst[at] = src[at];
    }
}

__global__ void k_peerSignal(uint32_t* remoteFlag, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(remoteFlag), "r"(value) : "memory");
}

__global__ void k_peerWait(uint32_t* localFlags, uint32_t flagIndex, uint32_t value) {
    const long long t0 = clock64();
    while (true) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(localFlags + flagIndex) : "memory");
        if ((int32_t)(v - value) >= 0)
            return;
        if (clock64() - t0 > 6000000000ll) { // ~3 s: the neighbour is gone; do not hang the GPU
            atomicExch(localFlags + (kPeerFlagWords - 1), 1u);
            return;
        }
        __nanosleep(200);
    }
}

// NCCL entry points resolved at run time, from the NCCL the host process has ALREADY loaded (the ncclComm_t handed to the library
// comes from it): the global scope first, then the loaded libnccl.so.2 by name with RTLD_NOLOAD (torch loads its bundled copy with
// local visibility).  Never loads a library: another libnccl of the same soname mapped into the process would shadow the
// host's own one for everything loaded later.
void* ncclSymbol(const char* name) {
    void* sym = dlsym(RTLD_DEFAULT, name);
    if (!sym) {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (lib)
            sym = dlsym(lib, name);
    }
    return sym;
}
} // namespace gfx

using namespace gfx;

extern "C" {

void* gfx_buffer_device_ptr(gfx_ctx* ctx, int bufferId, uint32_t index, size_t* bytes);

static int peerEnsure(gfx_ctx* ctx, PeerState** out) {
    PeerState &P = g_peers[ctx];
    if (!P.localFlags) {
        GFX_CUDA(ctx, cudaMalloc(&P.localFlags, kPeerFlagWords * 4));
        GFX_CUDA(ctx, cudaMemset(P.localFlags, 0, kPeerFlagWords * 4));
    }
    *out = &P;
    return GFX_OK;
}

int gfx_peer_export(gfx_ctx* ctx, int bufferId, uint32_t index, void* handle64) {
    if (!ctx || !handle64)
        return GFX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    PeerState* P;
    if (int rc = peerEnsure(ctx, &P))
        return rc;
    void* p = P->localFlags;
    if (bufferId != kPeerFlagsBufferId) {
        size_t bytes = 0;
        p = gfx_buffer_device_ptr(ctx, bufferId, index, &bytes);
        if (!p) {
            ctx->setError("gfx_peer_export: no such frame buffer");
            return GFX_ERR_INVALID_ARGUMENT;
        }
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    GFX_CUDA(ctx, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p));
    return GFX_OK;
}

int gfx_peer_open(gfx_ctx* ctx, uint32_t link, int bufferId, uint32_t index, const void* handle64) {
    if (!ctx || !handle64 || link > 1)
        return GFX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    PeerState* P;
    if (int rc = peerEnsure(ctx, &P))
        return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    GFX_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    P->links[link].mapped.push_back(p);
    if (bufferId == kPeerFlagsBufferId)
        P->links[link].flags = static_cast<uint32_t*>(p);
    else
        P->links[link].buffers[{ bufferId, index }] = p;
    return GFX_OK;
}

int gfx_peer_push_rows(gfx_ctx* ctx, void* stream, uint32_t link, int bufferId, uint32_t index, uint32_t rowLo, uint32_t rowHi) {
    if (!ctx || link > 1 || !ctx->frame.created)
        return GFX_ERR_INVALID_ARGUMENT;
    if (rowHi <= rowLo)
        return GFX_OK;
    PeerState &P = g_peers[ctx];
    const auto it = P.links[link].buffers.find({ bufferId, index });
    size_t bytes = 0;
    void* local = gfx_buffer_device_ptr(ctx, bufferId, index, &bytes);
    if (it == P.links[link].buffers.end() || !local || rowHi > ctx->frame.H) {
        ctx->setError("gfx_peer_push_rows: buffer not opened on this link or rows out of range");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    const size_t pixels = (size_t)ctx->frame.W * ctx->frame.H;
    const uint32_t planes = bufferId == GFX_BUF_RESERVOIR ? 3u : 1u; // SoA planes of 16 B (see abi.BUFFER_LAYOUT)
    const size_t bytesPerPixel = bytes / pixels / planes;
    if (bytesPerPixel % 8 != 0 || bytesPerPixel * pixels * planes != bytes) {
        ctx->setError("gfx_peer_push_rows: unsupported buffer layout");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    const size_t perPixel = bytesPerPixel / 8;
    const size_t planeStride = pixels * perPixel, rowOffset = (size_t)rowLo * ctx->frame.W * perPixel;
    const size_t count = (size_t)(rowHi - rowLo) * ctx->frame.W * perPixel;
    const uint32_t blocks = (uint32_t)std::min<size_t>((count * planes + 255) / 256, 148 * 8);
    cudaStream_t s = (cudaStream_t)stream;
    { GFX_TIMED(ctx, s, "peer_push_rows");
    k_peerPushRows<<<blocks, 256, 0, s>>>(static_cast<uint2*>(it->second), static_cast<const uint2*>(local), planes, planeStride,
                                          rowOffset, count); }
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

int gfx_peer_signal(gfx_ctx* ctx, void* stream, uint32_t link, uint32_t flagIndex, uint32_t value) {
    if (!ctx || link > 1 || flagIndex >= kPeerFlagWords - 1)
        return GFX_ERR_INVALID_ARGUMENT;
    PeerState &P = g_peers[ctx];
    if (!P.links[link].flags) {
        ctx->setError("gfx_peer_signal: link has no flag block");
        return GFX_ERR_NOT_READY;
    }
    k_peerSignal<<<1, 1, 0, (cudaStream_t)stream>>>(P.links[link].flags + flagIndex, value);
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

int gfx_peer_wait(gfx_ctx* ctx, void* stream, uint32_t flagIndex, uint32_t value) {
    if (!ctx || flagIndex >= kPeerFlagWords - 1)
        return GFX_ERR_INVALID_ARGUMENT;
    PeerState* P;
    if (int rc = peerEnsure(ctx, &P))
        return rc;
    k_peerWait<<<1, 1, 0, (cudaStream_t)stream>>>(P->localFlags, flagIndex, value);
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

int gfx_peer_status(gfx_ctx* ctx, void* stream, uint32_t* timedOut) {
    if (!ctx || !timedOut)
        return GFX_ERR_INVALID_ARGUMENT;
    PeerState* P;
    if (int rc = peerEnsure(ctx, &P))
        return rc;
    GFX_CUDA(ctx, cudaMemcpyAsync(timedOut, P->localFlags + (kPeerFlagWords - 1), 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GFX_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return GFX_OK;
}

typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*NcclCommUserRankFn)(void*, int*);

int gfx_framebuffer_allgather(gfx_ctx* ctx, void* ncclComm, void* stream, uint32_t rowsPerRank, void* dstFramebuffer) {
    if (!ctx || !ncclComm || !dstFramebuffer || !ctx->frame.created || rowsPerRank == 0)
        return GFX_ERR_INVALID_ARGUMENT;
    const NcclAllGatherFn allGather = reinterpret_cast<NcclAllGatherFn>(ncclSymbol("ncclAllGather"));
    const NcclCommUserRankFn userRank = reinterpret_cast<NcclCommUserRankFn>(ncclSymbol("ncclCommUserRank"));
    if (!allGather || !userRank) {
        ctx->setError("gfx_framebuffer_allgather: ncclAllGather not found (load NCCL in the host process)");
        return GFX_ERR_UNSUPPORTED;
    }
    // the rank's own strip inside its beauty buffer: NCCL's in-place convention is sendbuff = recvbuff + rank * count, which the
    // caller gets by passing the beauty buffer itself as dstFramebuffer; rank is implied by the communicator, so the send
    // pointer is computed from the communicator-independent row layout only when dst is the beauty buffer
    int rank = 0;
    if (userRank(ncclComm, &rank) != 0) {
        ctx->setError("gfx_framebuffer_allgather: ncclCommUserRank failed");
        return GFX_ERR_UNSUPPORTED;
    }
    const size_t count = (size_t)rowsPerRank * ctx->frame.W * 4; // floats per strip
    if ((size_t)(rank + 1) * rowsPerRank > ctx->frame.H)
        return GFX_ERR_INVALID_ARGUMENT;
    const float* send = reinterpret_cast<const float*>(ctx->frame.beauty) + (size_t)rank * count;
    const int rc = allGather(send, dstFramebuffer, count, 7 /* ncclFloat32 */, ncclComm, (cudaStream_t)stream);
    if (rc != 0) {
        ctx->setError("gfx_framebuffer_allgather: ncclAllGather returned " + std::to_string(rc));
        return GFX_ERR_CUDA;
    }
    return GFX_OK;
}

int gfx_peer_close(gfx_ctx* ctx) {
    if (!ctx)
        return GFX_ERR_INVALID_ARGUMENT;
    const auto it = g_peers.find(ctx);
    if (it == g_peers.end())
        return GFX_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (PeerLink &l : it->second.links)
        for (void* p : l.mapped)
            cudaIpcCloseMemHandle(p);
    cudaFree(it->second.localFlags);
    g_peers.erase(it);
    return GFX_OK;
}

int gfx_framebuffer_allgatherv(gfx_ctx* ctx, void* ncclComm, void* stream, const uint32_t* rowStarts, uint32_t world, void* dstFramebuffer) {
    if (!ctx || !ncclComm || !dstFramebuffer || !rowStarts || !ctx->frame.created || world == 0 || world > 64)
        return GFX_ERR_INVALID_ARGUMENT;
    typedef int (*NcclBroadcastFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    typedef int (*NcclGroupFn)();
    const NcclBroadcastFn broadcast = reinterpret_cast<NcclBroadcastFn>(ncclSymbol("ncclBroadcast"));
    const NcclGroupFn groupStart = reinterpret_cast<NcclGroupFn>(ncclSymbol("ncclGroupStart"));
    const NcclGroupFn groupEnd = reinterpret_cast<NcclGroupFn>(ncclSymbol("ncclGroupEnd"));
    if (!broadcast || !groupStart || !groupEnd) {
        ctx->setError("gfx_framebuffer_allgatherv: NCCL not found in the host process");
        return GFX_ERR_UNSUPPORTED;
    }
    const uint32_t H = ctx->frame.H;
    for (uint32_t r = 0; r < world; ++r)
        if (rowStarts[r] > rowStarts[r + 1] || rowStarts[r + 1] > H)
            return GFX_ERR_INVALID_ARGUMENT;
    // strips of unequal height: one broadcast per strip from its owner, fused into one NCCL launch by the group
    const size_t rowFloats = (size_t)ctx->frame.W * 4;
    const float* beauty = reinterpret_cast<const float*>(ctx->frame.beauty);
    float* dst = reinterpret_cast<float*>(dstFramebuffer);
    int rc = groupStart();
    for (uint32_t r = 0; r < world && rc == 0; ++r) {
        const size_t count = (size_t)(rowStarts[r + 1] - rowStarts[r]) * rowFloats;
        if (count)
            rc = broadcast(beauty + rowStarts[r] * rowFloats, dst + rowStarts[r] * rowFloats, count, 7 /* ncclFloat32 */, (int)r, ncclComm,
                           (cudaStream_t)stream);
    }
    const int rcEnd = groupEnd();
    if (rc != 0 || rcEnd != 0) {
        ctx->setError("gfx_framebuffer_allgatherv: NCCL returned " + std::to_string(rc ? rc : rcEnd));
        return GFX_ERR_CUDA;
    }
    return GFX_OK;
}

} // extern "C"
