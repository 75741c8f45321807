// present.cu — the output side of a frame: accumulation buffer -> tone-mapped, sRGB-encoded RGBA8.
//
// Replaces, in one kernel, copyToLinearBuffers + visualizeToOutputBuffer (restir_di/gpu_kernels/copy_buffers.cu:6-28,32-80),
// the display shader common/shaders/drawOptiXResult.frag with GL_FRAMEBUFFER_SRGB, and the arithmetic of the screenshot
// writer saveImage(float4*, SDRImageSaverConfig) (common/common_host.cpp:2859-2897), whose order of operations is followed:
// tone map, gamma, pack.  exp / pow come from detmath.h so that the oracle produces the same codes bit for bit.
// HBM-bound: 16 B read + 4 B written per pixel.
#include "scene.cuh"
#include "context.h"
#include "detmath.h"

namespace gfx {

GFX_D float srgbGamma(float value) { // sRGB_gamma_s, basic_types.h:5405-5410
    if (value <= 0.0031308f)
        return 12.92f * value;
    return 1.055f * dm_pow(value, 1.0f / 2.4f) - 0.055f;
}

__global__ void k_present(const float4* __restrict__ src, uint32_t* __restrict__ dst, uint32_t W, uint32_t H, int mode,
                          uint32_t flags, float brightnessScale, float alphaForOverride) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H)
        return;
    const uint32_t sy = (flags & GFX_PRESENT_FLIP_Y) ? H - 1 - y : y;
    float4 v = src[(size_t)sy * W + x];
    if (mode == GFX_PRESENT_NORMAL) { // copy_buffers.cu:22-25 then :59-66
        f3 n(v.x, v.y, v.z);
        if (n.x != 0 || n.y != 0 || n.z != 0)
            n = normalize(n);
        v = make_float4(0.5f + 0.5f * n.x, 0.5f + 0.5f * n.y, 0.5f + 0.5f * n.z, 1.0f);
    }
    if (alphaForOverride >= 0.0f)
        v.w = alphaForOverride;
    if (flags & GFX_PRESENT_TONE_MAP) {
        float r = v.x, g = v.y, b = v.z;
        if (!(isfinite(r) && isfinite(g) && isfinite(b)))
            r = g = b = 0.0f;
        const float lum = 0.2126729f * r + 0.7151522f * g + 0.0721750f * b;
        const float lumT = 1 - dm_exp(-(brightnessScale * lum)); // simpleToneMap_s
        const float s = lum > 0.0f ? lumT / lum : 0.0f;
        v.x = r * s;
        v.y = g * s;
        v.z = b * s;
    }
    if (flags & GFX_PRESENT_SRGB_GAMMA) {
        v.x = srgbGamma(v.x);
        v.y = srgbGamma(v.y);
        v.z = srgbGamma(v.z);
    }
    dst[(size_t)y * W + x] = (min(dm_f2uint(v.x * 255), 255u) << 0) | (min(dm_f2uint(v.y * 255), 255u) << 8) |
                             (min(dm_f2uint(v.z * 255), 255u) << 16) | (min(dm_f2uint(v.w * 255), 255u) << 24);
}

} // namespace gfx

using namespace gfx;

extern "C" void* gfx_buffer_device_ptr(gfx_ctx* ctx, int bufferId, uint32_t index, size_t* bytes);

namespace gfx {

int launchPresent(gfx_ctx* ctx, cudaStream_t stream, const GfxPresentParams* p) {
    FrameState &F = ctx->frame;
    const size_t n = (size_t)F.W * F.H;
    if (p->sourceBuffer == GFX_BUF_PRESENT_RGBA8 || (p->mode != GFX_PRESENT_COLOR && p->mode != GFX_PRESENT_NORMAL)) {
        ctx->setError("gfx_present_launch: bad source buffer or mode");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    size_t bytes = 0;
    const void* src = gfx_buffer_device_ptr(ctx, p->sourceBuffer, p->sourceIndex, &bytes);
    if (!src || bytes != n * 16) {
        ctx->setError("gfx_present_launch: the source must be a float4-per-pixel frame buffer");
        return GFX_ERR_INVALID_ARGUMENT;
    }
    if (!F.presentRgba8)
        GFX_CUDA(ctx, cudaMalloc(&F.presentRgba8, n * 4));
    const dim3 block(32, 8), grid((F.W + 31) / 32, (F.H + 7) / 8);
    { GFX_TIMED(ctx, stream, "present");
    k_present<<<grid, block, 0, stream>>>(static_cast<const float4*>(src), F.presentRgba8, F.W, F.H, p->mode, p->flags,
                                          p->brightnessScale, p->alphaForOverride); }
    ctx->launches++;
    GFX_CUDA(ctx, cudaGetLastError());
    return GFX_OK;
}

} // namespace gfx
