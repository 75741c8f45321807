// oracle/present.cpp — TEST INFRASTRUCTURE.  CPU restatement of the reference's output side for one image:
//   copyToLinearBuffers / visualizeToOutputBuffer  restir_di/gpu_kernels/copy_buffers.cu:6-28, 32-80
//   saveImage(float4*, SDRImageSaverConfig)        common/common_host.cpp:2859-2897
//   simpleToneMap_s / sRGB_gamma_s                 common/basic_types.h:5391-5410
//   sRGB_calcLuminance                             common/shaders/drawOptiXResult.frag:12-14
// exp / pow are detmath.h's (shared with the CUDA kernel; libm and libdevice differ in the last bits, which would flip
// 8-bit codes); float -> uint conversion saturates like CUDA's.  Only tests/, smoke() and bench.py's cpu_baseline leg may
// call this.  Parity unpinned against the reference binary (see oracle.h).
#include "oracle.h"
#include "vecmath.h"
#include "../gfxexp_b200/csrc/detmath.h"
#include <cmath>
#include <algorithm>
using namespace orc;

static inline float simpleToneMap(float value) { return 1 - dm_exp(-value); }          // :5391-5394
static inline float srgbGamma(float value) {                                            // :5405-5410
    if (value <= 0.0031308f)
        return 12.92f * value;
    return 1.055f * dm_pow(value, 1 / 2.4f) - 0.055f;
}

extern "C" void orc_present(const float* srcRGBA, uint32_t width, uint32_t height, const GfxPresentParams* config, uint32_t* image) {
    for (uint32_t y = 0; y < height; ++y) {
        const uint32_t sy = (config->flags & GFX_PRESENT_FLIP_Y) ? (height - 1 - y) : y; // :2865
        for (uint32_t x = 0; x < width; ++x) {
            const float* s = srcRGBA + 4 * ((size_t)sy * width + x);
            float src[4] = { s[0], s[1], s[2], s[3] };
            if (config->mode == GFX_PRESENT_NORMAL) {
                float3 normal(src[0], src[1], src[2]);                                   // copy_buffers.cu:22-25
                if (normal.x != 0 || normal.y != 0 || normal.z != 0)
                    normal = normalize(normal);
                src[0] = 0.5f + 0.5f * normal.x;                                         // copy_buffers.cu:62-64
                src[1] = 0.5f + 0.5f * normal.y;
                src[2] = 0.5f + 0.5f * normal.z;
                src[3] = 1.0f;
            }
            if (config->alphaForOverride >= 0.0f)                                        // :2868-2869
                src[3] = config->alphaForOverride;
            if (config->flags & GFX_PRESENT_TONE_MAP) {                                  // :2870-2880
                float r = src[0], g = src[1], b = src[2];
                if (!(std::isfinite(r) && std::isfinite(g) && std::isfinite(b)))
                    r = g = b = 0.0f;
                const float lum = 0.2126729f * r + 0.7151522f * g + 0.0721750f * b;
                const float lumT = simpleToneMap(config->brightnessScale * lum);
                const float scale = lum > 0.0f ? lumT / lum : 0.0f;
                src[0] = r * scale;
                src[1] = g * scale;
                src[2] = b * scale;
            }
            if (config->flags & GFX_PRESENT_SRGB_GAMMA) {                                // :2881-2885
                src[0] = srgbGamma(src[0]);
                src[1] = srgbGamma(src[1]);
                src[2] = srgbGamma(src[2]);
            }
            image[(size_t)y * width + x] =                                               // :2886-2890
                (std::min<uint32_t>(dm_f2uint(src[0] * 255), 255) << 0) | (std::min<uint32_t>(dm_f2uint(src[1] * 255), 255) << 8) |
                (std::min<uint32_t>(dm_f2uint(src[2] * 255), 255) << 16) | (std::min<uint32_t>(dm_f2uint(src[3] * 255), 255) << 24);
        }
    }
}
