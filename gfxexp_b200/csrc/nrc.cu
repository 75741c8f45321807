// nrc.cu — NRC network (placeholder until the tcgen05 MLP lands; see DESIGN.md build order).
#include "context.h"
struct gfx_nrc { gfx_ctx* ctx; };
extern "C" {
int gfx_nrc_create(gfx_ctx* ctx, uint32_t, float, gfx_nrc** out) { if (out) *out = nullptr; if (ctx) ctx->setError("gfx_nrc_create: not implemented yet"); return GFX_ERR_UNSUPPORTED; }
void gfx_nrc_destroy(gfx_nrc* nrc) { delete nrc; }
int gfx_nrc_infer(gfx_nrc*, void*, const float*, float*, uint32_t) { return GFX_ERR_UNSUPPORTED; }
int gfx_nrc_train(gfx_nrc*, void*, const float*, const float*, uint32_t, float*) { return GFX_ERR_UNSUPPORTED; }
int gfx_nrc_get_params(gfx_nrc*, void*, size_t) { return GFX_ERR_UNSUPPORTED; }
int gfx_nrc_set_params(gfx_nrc*, const void*, size_t) { return GFX_ERR_UNSUPPORTED; }
uint32_t gfx_nrc_num_params(gfx_nrc*) { return 0; }
}
