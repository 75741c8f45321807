"""GPU parity: the rearchitected ReSTIR DI renderer (light presampling, per-pixel RIS from a per-tile subset, decoupled
shadow rays with reusable visibility bits, pairwise-MIS shading + resampling) vs the CPU oracle, bit-exact
(SURVEY.md §8a row R5)."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, scenes

pytestmark = pytest.mark.gpu

BUFFERS = [(abi.BUF_RNG, 1), (abi.BUF_RESERVOIR, 2), (abi.BUF_RESERVOIR_INFO, 2), (abi.BUF_SAMPLE_VISIBILITY, 2),
           (abi.BUF_BEAUTY_ACCUM, 1)]


def _same(got, want, tag):
    g, w = np.ascontiguousarray(got), np.ascontiguousarray(want)
    g = g.view(np.uint32) if g.dtype.itemsize == 4 else g
    w = w.view(np.uint32) if w.dtype.itemsize == 4 else w
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        raise AssertionError(f"{tag}: {len(bad)} elements differ, first {bad[:4].tolist()}: "
                             f"{got[tuple(bad[0])]} vs {want[tuple(bad[0])]}")


def _run(gfx_ctx, oracle, scene, w, h, frames, configure, move_camera=False):
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    gfx_ctx.create_frame(w, h)
    oscene = oracle.OracleScene(scene)
    oframe = oracle.OracleFrame(oscene, w, h)
    p = abi.default_frame_params(scene, w, h)
    kwargs = configure(p)
    rays = 0
    for frame in range(frames):
        p.numAccumFrames = frame
        if move_camera and frame > 0:
            p.prevCamera = p.camera
            cam = abi.GfxCamera()
            C = __import__("ctypes")
            C.memmove(C.byref(cam), C.byref(p.camera), C.sizeof(cam))
            cam.position[0] += 0.35
            cam.position[2] -= 0.2
            p.camera = cam
        gfx_ctx.build_light_distributions(frame % 2)
        for kind, pass_id in engine.restir_rearch_frame_passes(p, frame, **kwargs):
            if kind == "gbuffer":
                gfx_ctx.gbuffer(p)
                oframe.gbuffer(p)
            else:
                gfx_ctx.restir(p, pass_id)
                rays += oframe.restir_rearch(p, pass_id)
                if pass_id == abi.RESTIR_PRESAMPLE_LIGHTS:
                    _same(gfx_ctx.download_linear(abi.BUF_PRESAMPLED_LIGHTS), oframe.linear_buffer(abi.BUF_PRESAMPLED_LIGHTS),
                          f"frame {frame} presampled lights")
                for buf, count in BUFFERS:
                    for idx in range(count):
                        _same(gfx_ctx.download(buf, idx), oframe.buffer(buf, idx), f"frame {frame} pass {pass_id} buffer {buf}[{idx}]")
    beauty = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3]
    assert np.isfinite(beauty).all() and beauty.mean() > 1e-3
    assert rays > 0
    return rays / (frames * w * h)


def test_rearch_biased_spatiotemporal(gfx_ctx, oracle):
    """the default configuration (restir_di_main.cpp:1938-1966): temporal + spatiotemporal reuse, reused temporal visibility"""
    rays = _run(gfx_ctx, oracle, scenes.small_city_scene(), 192, 108, 4, lambda p: dict(temporal=True, spatial=True, unbiased=False))
    assert rays < 3.0


@pytest.mark.parametrize("unbiased", [False, True])
def test_rearch_with_environment_light(gfx_ctx, oracle, unbiased):
    """the first quarter of every presampled light subset comes from the environment's importance map (per_pixel_ris.cu:12-28);
    miss pixels show the environment"""
    _run(gfx_ctx, oracle, scenes.small_city_scene_env(), 160, 96, 3, lambda p: dict(temporal=True, spatial=True, unbiased=unbiased))


def test_rearch_with_image_textures(gfx_ctx, oracle):
    _run(gfx_ctx, oracle, scenes.small_city_scene_textured(), 160, 96, 3, lambda p: dict(temporal=True, spatial=True, unbiased=False))


def test_rearch_unbiased_moving_camera_random_neighbors(gfx_ctx, oracle):
    def configure(p):
        p.useLowDiscrepancyNeighbors = 0
        p.enableJittering = 1
        return dict(temporal=True, spatial=True, unbiased=True)
    rays = _run(gfx_ctx, oracle, scenes.small_city_scene(), 160, 96, 4, configure, move_camera=True)
    assert rays > 3.0  # up to 9 visibility tests per pixel in the reference, 7 distinct rays


@pytest.mark.parametrize("temporal,spatial", [(True, False), (False, True), (False, False)])
def test_rearch_partial_reuse_and_spatial_visibility_reuse(gfx_ctx, oracle, temporal, spatial):
    def configure(p):
        p.reuseVisibilityForSpatiotemporal = 1
        p.radiusThresholdForSpatialVisReuse = 12.0
        p.log2NumCandidateSamples = 3
        return dict(temporal=temporal, spatial=spatial, unbiased=False)
    _run(gfx_ctx, oracle, scenes.tiny_city_scene(), 100, 60, 3, configure)
