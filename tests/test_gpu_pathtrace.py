"""GPU parity: the wavefront path tracer vs the CPU oracle's per-pixel megakernel restatement, bit-exact
(NEE + MIS + Russian roulette, path_tracing/gpu_kernels/optix_pathtracing_kernels.cu)."""
import numpy as np
import pytest

from gfxexp_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def _setup(ctx, oracle, scene, w, h):
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(w, h)
    oscene = oracle.OracleScene(scene)
    oframe = oracle.OracleFrame(oscene, w, h)
    return oscene, oframe


def _assert_same(got, want, tag):
    g = got.view(np.uint32) if got.dtype != np.uint64 else got
    w = want.view(np.uint32) if want.dtype != np.uint64 else want
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        raise AssertionError(f"{tag}: {len(bad)} elements differ, first {bad[:4].tolist()}: "
                             f"{got[tuple(bad[0][:2])]} vs {want[tuple(bad[0][:2])]}")


@pytest.mark.parametrize("max_path_length,scene_name", [(2, "small_city_scene"), (5, "small_city_scene"), (9, "small_city_scene"),
                                                        (5, "small_interior_scene"),
                                                        (5, "small_city_scene_env"), (4, "env_only_scene"),
                                                        (5, "small_city_scene_textured")])
def test_pathtrace_accumulation_bit_exact(gfx_ctx, oracle, max_path_length, scene_name):
    """three accumulated samples per pixel in a closed scene with ~100 emitters: every pixel's radiance and RNG
    state equal the oracle's, i.e. the wavefront reordering changes neither the draws nor the summation order
    (small_interior_scene: SimplePBR materials, many two-triangle emitters in a closed room)"""
    scene = getattr(scenes, scene_name)()
    w, h = 160, 90
    _, oframe = _setup(gfx_ctx, oracle, scene, w, h)
    p = abi.default_frame_params(scene, w, h)
    p.maxPathLength = max_path_length
    gfx_ctx.build_light_distributions(0)
    rays = 0
    for frame in range(3):
        p.numAccumFrames = frame
        p.frameIndex = frame
        gfx_ctx.gbuffer(p)
        gfx_ctx.pathtrace(p)
        oframe.gbuffer(p)
        rays = oframe.pathtrace(p)
        _assert_same(gfx_ctx.download(abi.BUF_RNG), oframe.buffer(abi.BUF_RNG), f"rng frame {frame}")
        _assert_same(gfx_ctx.download(abi.BUF_BEAUTY_ACCUM), oframe.buffer(abi.BUF_BEAUTY_ACCUM), f"beauty frame {frame}")
    beauty = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3]
    assert np.isfinite(beauty).all() and beauty.mean() > 1e-3
    assert rays > w * h  # more than one ray per pixel: paths really bounce


def test_pathtrace_config1_like(gfx_ctx, oracle):
    """config 1 of BASELINE.json (512x512 single object under one rectangle light, maxPathLength 5) at reduced
    resolution, with jittered primary rays: misses return the 0.001 background, hits match bit for bit"""
    scene = scenes.teapot_like_scene()
    w = h = 128
    _, oframe = _setup(gfx_ctx, oracle, scene, w, h)
    p = abi.default_frame_params(scene, w, h)
    p.enableJittering = 1
    gfx_ctx.build_light_distributions(0)
    for frame in range(2):
        p.numAccumFrames = frame
        gfx_ctx.gbuffer(p)
        gfx_ctx.pathtrace(p)
        oframe.gbuffer(p)
        oframe.pathtrace(p)
    _assert_same(gfx_ctx.download(abi.BUF_RNG), oframe.buffer(abi.BUF_RNG), "rng")
    _assert_same(gfx_ctx.download(abi.BUF_BEAUTY_ACCUM), oframe.buffer(abi.BUF_BEAUTY_ACCUM), "beauty")
    beauty = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3]
    # the corner pixels never see the object: both samples are the 0.001 background
    assert np.all(beauty[0, 0] == np.float32(0.001)) and np.all(beauty[-1, -1] == np.float32(0.001))


def test_pathtrace_strip_equals_full_frame(gfx_ctx, oracle):
    """multi-GPU sharding contract: tracing the frame as two row strips gives the full-frame image"""
    scene = scenes.tiny_city_scene()
    w, h = 96, 64
    _, oframe = _setup(gfx_ctx, oracle, scene, w, h)
    p = abi.default_frame_params(scene, w, h)
    gfx_ctx.build_light_distributions(0)
    gfx_ctx.gbuffer(p)
    for y0, rows in ((0, 40), (40, 24)):
        p.tileOriginY, p.tileRows = y0, rows
        gfx_ctx.pathtrace(p)
    p.tileOriginY, p.tileRows = 0, 0
    oframe.gbuffer(p)
    oframe.pathtrace(p)
    _assert_same(gfx_ctx.download(abi.BUF_BEAUTY_ACCUM), oframe.buffer(abi.BUF_BEAUTY_ACCUM), "beauty")
