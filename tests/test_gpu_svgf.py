"""GPU parity: SVGF (temporal accumulation, variance, 5 a-trous stages, background, albedo modulation +
TAA) on ReSTIR DI output vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, scenes

pytestmark = pytest.mark.gpu

SVGF_BUFFERS = [(abi.BUF_SVGF_LIGHTING_VARIANCE, 2), (abi.BUF_SVGF_MOMENTS, 2), (abi.BUF_SVGF_PREV_LIGHTING, 1),
                (abi.BUF_SVGF_ALBEDO, 1), (abi.BUF_SVGF_DEPTH, 2), (abi.BUF_SVGF_FINAL, 2)]


def _run(gfx_ctx, oracle, scene, w, h, frames, pan, env_rotation=0.0):
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    gfx_ctx.create_frame(w, h)
    oscene = oracle.OracleScene(scene)
    oframe = oracle.OracleFrame(oscene, w, h)
    osvgf = oracle.OracleSvgf(oframe)
    p = abi.default_frame_params(scene, w, h)
    p.envLightRotation = env_rotation
    for frame in range(frames):
        cam = abi.make_camera(scene, w, h)
        cam.position[0] += pan * frame
        p.prevCamera = p.camera if frame > 0 else cam
        p.camera = cam
        p.numAccumFrames = 0  # SVGF consumes the current frame's 1-spp lighting
        gfx_ctx.build_light_distributions(frame % 2)
        for kind, pass_id in engine.restir_frame_passes(p, frame, 1):
            if kind == "gbuffer":
                gfx_ctx.gbuffer(p)
                oframe.gbuffer(p)
            else:
                gfx_ctx.restir(p, pass_id)
                oframe.restir(p, pass_id)
        for pass_id, stage in engine.svgf_frame_passes(p, frame):
            gfx_ctx.svgf(p, pass_id, stage)
            osvgf.run(p, pass_id, stage)
        gfx_ctx.synchronize()
        for buf, count in SVGF_BUFFERS:
            for idx in range(count):
                got = gfx_ctx.download(buf, idx).view(np.uint32)
                want = osvgf.buffer(buf, idx).view(np.uint32)
                assert np.array_equal(got, want), \
                    f"frame {frame}: SVGF buffer {buf}[{idx}] differs at {np.count_nonzero(got != want)} elements"
    return gfx_ctx.download(abi.BUF_SVGF_FINAL, (frames - 1) % 2), gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)


def test_svgf_static_camera_bit_exact(gfx_ctx, oracle):
    scene = scenes.small_city_scene()
    w, h = 192, 108
    final, noisy = _run(gfx_ctx, oracle, scene, w, h, 4, 0.0)
    assert np.isfinite(final).all()
    # the filter must actually denoise: closer to a converged render than the 1-spp input.
    # (GPU only: accumulate 96 more ReSTIR frames into the beauty buffer as the reference image.)
    p = abi.default_frame_params(scene, w, h)
    for k in range(96):
        p.numAccumFrames = k
        for kind, pass_id in engine.restir_frame_passes(p, 4 + k, 1):
            if kind == "gbuffer":
                gfx_ctx.gbuffer(p)
            else:
                gfx_ctx.restir(p, pass_id)
        p.numAccumFrames = k
    gfx_ctx.synchronize()
    converged = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3]

    def mse(img):
        return float(np.mean((np.clip(img[..., :3], 0, 4) - np.clip(converged, 0, 4)) ** 2))
    assert mse(final) < 0.5 * mse(noisy), (mse(final), mse(noisy))


def test_svgf_panning_camera_bit_exact(gfx_ctx, oracle):
    # ~2 px/frame pan exercises the bilinear reprojection and the disocclusion tests
    final, _ = _run(gfx_ctx, oracle, scenes.tiny_city_scene(), 160, 90, 3, 0.06)
    assert np.isfinite(final).all()


def test_svgf_with_environment_light(gfx_ctx, oracle):
    """miss pixels show the environment behind the filtered image (fillBackground, svgf.cu:431-437), the rotation quirk included;
    the ReSTIR frame under the filter samples the environment too"""
    final, _ = _run(gfx_ctx, oracle, scenes.small_city_scene_env(), 160, 90, 3, 0.04, env_rotation=0.7)
    assert np.isfinite(final).all()
    gb0 = gfx_ctx.download(abi.BUF_GBUFFER0, 0)
    sky = gb0[..., 0] == 0xFFFFFFFF
    assert sky.sum() > 100 and final[sky][:, :3].mean() > 0.05   # not the 0.001 background
