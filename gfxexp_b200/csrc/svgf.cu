// svgf.cu — SVGF passes (placeholder until the kernels land; see DESIGN.md build order).
#include "context.h"
namespace gfx {
int launchSVGF(gfx_ctx* ctx, cudaStream_t, const GfxFrameParams*, int, uint32_t) {
    ctx->setError("gfx_svgf_launch: not implemented yet");
    return GFX_ERR_UNSUPPORTED;
}
}
