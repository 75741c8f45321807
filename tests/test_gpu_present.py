"""GPU parity of the output side: gfx_present_launch (tone map + sRGB + RGBA8 packing, normal
visualisation) equals the oracle's restatement of saveImage / visualizeToOutputBuffer bit for bit, on a rendered frame."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, imageio, scenes

pytestmark = pytest.mark.gpu


def test_present_bit_exact_and_png(gfx_ctx, oracle, tmp_path):
    scene = scenes.tiny_city_scene()
    w, h = 160, 90
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    gfx_ctx.create_frame(w, h)
    p = abi.default_frame_params(scene, w, h)
    p.log2NumCandidateSamples = 3
    for frame in range(2):
        gfx_ctx.build_light_distributions(frame % 2)
        for kind, pass_id in engine.restir_frame_passes(p, frame, 1):
            gfx_ctx.gbuffer(p) if kind == "gbuffer" else gfx_ctx.restir(p, pass_id)
    gfx_ctx.synchronize()
    beauty = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)
    beauty[0, 0, :3] = np.nan                      # the saver turns non-finite colours into black
    beauty[0, 1, :3] = [np.inf, 0.5, 0.5]
    gfx_ctx.upload(abi.BUF_BEAUTY_ACCUM, 0, beauty)
    cases = [(abi.BUF_BEAUTY_ACCUM, beauty, abi.PRESENT_COLOR, abi.PRESENT_TONE_MAP | abi.PRESENT_SRGB_GAMMA, 2.0, 1.0),
             (abi.BUF_BEAUTY_ACCUM, beauty, abi.PRESENT_COLOR, abi.PRESENT_TONE_MAP | abi.PRESENT_SRGB_GAMMA | abi.PRESENT_FLIP_Y, 10.0, -1.0),
             (abi.BUF_BEAUTY_ACCUM, beauty, abi.PRESENT_COLOR, 0, 1.0, 0.5),
             (abi.BUF_ALBEDO_ACCUM, gfx_ctx.download(abi.BUF_ALBEDO_ACCUM), abi.PRESENT_COLOR, abi.PRESENT_SRGB_GAMMA, 1.0, 1.0),
             (abi.BUF_NORMAL_ACCUM, gfx_ctx.download(abi.BUF_NORMAL_ACCUM), abi.PRESENT_NORMAL, 0, 1.0, -1.0)]
    for buf, src, mode, flags, brightness, alpha in cases:
        got = gfx_ctx.present(buf, 0, mode, flags, brightness, alpha)
        want = oracle.present(src, mode, flags, brightness, alpha)
        assert got.shape == (h, w) and got.dtype == np.uint32
        bad = np.argwhere(got != want)
        assert len(bad) == 0, f"buffer {buf} mode {mode} flags {flags}: {len(bad)} pixels differ, first {bad[:3].tolist()}"
    img = gfx_ctx.present(abi.BUF_BEAUTY_ACCUM, 0, abi.PRESENT_COLOR, abi.PRESENT_TONE_MAP | abi.PRESENT_SRGB_GAMMA, 2.0, 1.0)
    assert (img & 0xFFFFFF).any() and (img >> 24 == 255).all()
    assert (img[0, 0] & 0xFFFFFF) == 0 and (img[0, 1] & 0xFFFFFF) == 0
    path = str(tmp_path / "frame.png")
    imageio.write_png(path, img)
    assert np.array_equal(imageio.read_png(path), img)
    with pytest.raises(engine.GfxError):
        gfx_ctx.present(abi.BUF_GBUFFER1, 0)        # not a float4-per-pixel buffer
