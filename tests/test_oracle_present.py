"""CPU checks of the output side: detmath's log / pow against libm, the oracle's tone map + sRGB packing against a float64
evaluation of the same formulas (within one 8-bit code), and the PNG writer's round trip."""
import ctypes as C

import numpy as np

from gfxexp_b200 import abi, imageio


def test_dm_log_and_pow(oracle):
    L = oracle.lib()
    x = np.concatenate([np.geomspace(1e-30, 1e30, 200001), np.linspace(0.5, 2.0, 100001)]).astype(np.float32)
    y = np.empty_like(x)
    L.orc_dm_log(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    ref = np.log(x.astype(np.float64))
    assert np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1.0)) < 2.5e-7
    v = np.linspace(0.0, 4.0, 400001).astype(np.float32)
    e = np.full_like(v, np.float32(1 / 2.4))
    p = np.empty_like(v)
    L.orc_dm_pow(v.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), C.c_size_t(v.size))
    refp = v.astype(np.float64) ** (1 / 2.4)
    assert p[0] == 0.0
    assert np.max(np.abs(p[1:] - refp[1:]) / refp[1:]) < 1e-6


def _present_f64(img, flags, brightness, alpha):
    v = img.astype(np.float64).copy()
    rgb = v[..., :3]
    if flags & abi.PRESENT_TONE_MAP:
        bad = ~np.isfinite(rgb).all(-1)
        rgb[bad] = 0.0
        lum = 0.2126729 * rgb[..., 0] + 0.7151522 * rgb[..., 1] + 0.0721750 * rgb[..., 2]
        s = np.where(lum > 0, (1 - np.exp(-brightness * lum)) / np.where(lum > 0, lum, 1.0), 0.0)
        rgb *= s[..., None]
    if flags & abi.PRESENT_SRGB_GAMMA:
        rgb[:] = np.where(rgb <= 0.0031308, 12.92 * rgb, 1.055 * np.maximum(rgb, 0) ** (1 / 2.4) - 0.055)
    v[..., 3] = alpha
    return np.clip(np.floor(v * 255), 0, 255).astype(np.int64)


def test_present_matches_float64_evaluation(oracle):
    rng = np.random.default_rng(5)
    img = (rng.gamma(0.7, 1.5, size=(40, 56, 4))).astype(np.float32)
    img[0, 0, :3] = np.nan
    img[0, 1, :3] = 0.0
    img[0, 2, :3] = [np.inf, 1.0, 1.0]
    for flags in (0, abi.PRESENT_TONE_MAP, abi.PRESENT_SRGB_GAMMA, abi.PRESENT_TONE_MAP | abi.PRESENT_SRGB_GAMMA):
        got = oracle.present(img, abi.PRESENT_COLOR, flags, 2.0, 1.0)
        chans = np.stack([(got >> s) & 0xFF for s in (0, 8, 16, 24)], axis=-1).astype(np.int64)
        finite = np.isfinite(img[..., :3]).all(-1)
        if flags & abi.PRESENT_TONE_MAP:
            want = _present_f64(img, flags, 2.0, 1.0)
            assert np.abs(chans[finite] - want[finite]).max() <= 1, flags
            assert (chans[0, 0, :3] == 0).all() and (chans[0, 2, :3] == 0).all()   # non-finite colours become black
        else:
            want = _present_f64(img, flags, 2.0, 1.0)
            assert np.abs(chans[finite] - want[finite]).max() <= 1, flags
        assert (chans[..., 3] == 255).all()
    flipped = oracle.present(img, abi.PRESENT_COLOR, abi.PRESENT_FLIP_Y | 3, 2.0, 1.0)
    assert np.array_equal(flipped[::-1], oracle.present(img, abi.PRESENT_COLOR, 3, 2.0, 1.0))
    normals = np.zeros((4, 4, 4), dtype=np.float32)
    normals[..., :3] = [0.0, 3.0, 4.0]
    normals[0, 0, :3] = 0.0
    n8 = oracle.present(normals, abi.PRESENT_NORMAL, 0, 1.0, -1.0)
    assert (n8[1, 1] & 0xFF, (n8[1, 1] >> 8) & 0xFF, (n8[1, 1] >> 16) & 0xFF) == (127, 204, 229)   # 0.5 + 0.5 * (0, .6, .8)
    assert (n8[0, 0] & 0xFFFFFF) == 0x7F7F7F


def test_png_round_trip(tmp_path, oracle):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 2 ** 32, size=(37, 53), dtype=np.uint32)
    path = str(tmp_path / "a.png")
    imageio.write_png(path, img)
    assert np.array_equal(imageio.read_png(path), img)
