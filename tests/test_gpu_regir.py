"""GPU parity: ReGIR cell reservoirs (build + temporal reuse + access bookkeeping) and the ReGIR path tracer vs the CPU
oracle, bit-exact."""
import numpy as np
import pytest

from gfxexp_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def _same(got, want, tag):
    g, w = np.ascontiguousarray(got), np.ascontiguousarray(want)
    g = g.view(np.uint32) if g.dtype.itemsize == 4 else g
    w = w.view(np.uint32) if w.dtype.itemsize == 4 else w
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        raise AssertionError(f"{tag}: {len(bad)} elements differ, first {bad[:4].tolist()}: "
                             f"{got[tuple(bad[0])]} vs {want[tuple(bad[0])]}")


@pytest.mark.parametrize("grid,randomize,scene_name", [((8, 4, 8), True, "small_city_scene"), ((5, 3, 6), False, "small_city_scene"),
                                                       ((8, 4, 8), True, "small_city_scene_env"),
                                                       ((5, 3, 6), True, "small_city_scene_textured")])
def test_regir_frames_bit_exact(gfx_ctx, oracle, grid, randomize, scene_name):
    # small_city_scene_env: the cell reservoirs also stream environment-light candidates (build_cell_reservoirs.cu:120-139)
    scene = getattr(scenes, scene_name)()
    w, h = 160, 90
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    gfx_ctx.create_frame(w, h)
    gfx_ctx.build_light_distributions(0)
    oscene = oracle.OracleScene(scene)
    oframe = oracle.OracleFrame(oscene, w, h)
    p = abi.default_frame_params(scene, w, h)
    p.regirGridDim = (abi.c_u32 * 3)(*grid)
    p.regirEnableCellRandomization = 1 if randomize else 0
    # frames 0..2 warm the grid up; from frame 9 on the cells nobody looked at for 8 frames go idle
    for frame in list(range(3)) + [9, 10]:
        p.frameIndex, p.bufferIndex, p.numAccumFrames = frame, frame % 2, min(frame, 3)
        temporal = frame > 0
        gfx_ctx.gbuffer(p)
        gfx_ctx.regir_build_cells(p, frame, temporal)
        oframe.gbuffer(p)
        oframe.regir_build_cells(p, frame, temporal)
        _same(gfx_ctx.download_linear(abi.BUF_REGIR_SLOTS, p.bufferIndex, params=p),
              oframe.linear_buffer(abi.BUF_REGIR_SLOTS, p.bufferIndex, params=p), f"frame {frame} slots")
        _same(gfx_ctx.download_linear(abi.BUF_REGIR_SLOT_RNG, params=p), oframe.linear_buffer(abi.BUF_REGIR_SLOT_RNG, params=p),
              f"frame {frame} slot rngs")
        gfx_ctx.pathtrace(p, abi.PT_REGIR)
        oframe.pathtrace(p, abi.PT_REGIR)
        _same(gfx_ctx.download(abi.BUF_RNG), oframe.buffer(abi.BUF_RNG), f"frame {frame} rng")
        _same(gfx_ctx.download(abi.BUF_BEAUTY_ACCUM), oframe.buffer(abi.BUF_BEAUTY_ACCUM), f"frame {frame} beauty")
        _same(gfx_ctx.download_linear(abi.BUF_REGIR_CELL_ACCESSES, params=p), oframe.linear_buffer(abi.BUF_REGIR_CELL_ACCESSES, params=p),
              f"frame {frame} accesses")
        gfx_ctx.regir_update_access(p, frame)
        oframe.regir_update_access(p, frame)
        _same(gfx_ctx.download_linear(abi.BUF_REGIR_LAST_ACCESS, params=p), oframe.linear_buffer(abi.BUF_REGIR_LAST_ACCESS, params=p),
              f"frame {frame} last access")
        _same(gfx_ctx.download_linear(abi.BUF_REGIR_NUM_ACTIVE_CELLS, params=p)[p.bufferIndex],
              oframe.linear_buffer(abi.BUF_REGIR_NUM_ACTIVE_CELLS, params=p)[p.bufferIndex], f"frame {frame} active cells")
    last = gfx_ctx.download_linear(abi.BUF_REGIR_LAST_ACCESS, params=p)[:, 0]
    assert (last == 10).any() and (last == 0xFFFFFFFF).any()  # touched cells and never-touched cells
    beauty = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3]
    assert np.isfinite(beauty).all() and beauty.mean() > 1e-3
