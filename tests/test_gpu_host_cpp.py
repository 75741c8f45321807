"""The headless C++ host (host/gfx_headless.cpp) drives the library through the C ABI alone and must produce, bit for bit,
the images of the Python host for every renderer - i.e. the boundary really is language neutral and the C++ frame loops
of INTEGRATION.md are the ones the parity tests exercise."""
import json
import os
import subprocess

import numpy as np
import pytest

from gfxexp_b200 import abi, engine, scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gfxexp_b200", "gfx_headless")
W, H, FRAMES = 96, 64, 3


def _python_host(ctx, scene, renderer):
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(W, H)
    p = abi.default_frame_params(scene, W, H)
    net = None
    if renderer == "nrc":
        net = engine.NeuralRadianceCache(ctx, 2, 1e-2)
        net.set_params(engine.random_nrc_params(net.num_params, 64 * 64 * 2 + 16 * 64, grid_amplitude=0.1))
        rs = np.random.RandomState(72139121)  # std::mt19937 perFrameRng(72139121)
    for f in range(FRAMES):
        p.numAccumFrames = f
        ctx.build_light_distributions(f % 2)
        if renderer in ("restir", "restir_unbiased"):
            for kind, pid in engine.restir_frame_passes(p, f, 1, True, renderer.endswith("unbiased")):
                ctx.gbuffer(p) if kind == "gbuffer" else ctx.restir(p, pid)
        elif renderer in ("rearch", "rearch_unbiased"):
            for kind, pid in engine.restir_rearch_frame_passes(p, f, True, True, renderer.endswith("unbiased")):
                ctx.gbuffer(p) if kind == "gbuffer" else ctx.restir(p, pid)
        elif renderer == "pathtrace":
            p.frameIndex, p.bufferIndex = f, f % 2
            ctx.gbuffer(p)
            ctx.pathtrace(p)
        elif renderer == "regir":
            ctx.regir_frame(p, f)
        elif renderer == "nrc":
            offsets = [int(v) for v in rs.randint(0, 2 ** 32, size=2, dtype=np.uint64)]
            ctx.nrc_frame(net, p, f, offsets, train=True)
    out = ctx.download(abi.BUF_BEAUTY_ACCUM)
    if net is not None:
        net.close()
    return out


@pytest.mark.parametrize("renderer", ["restir", "restir_unbiased", "rearch", "rearch_unbiased", "pathtrace", "regir", "nrc"])
def test_cpp_host_matches_python_host(gfx_ctx, tmp_path, renderer):
    assert os.path.exists(HOST), "build the C++ host first (__graft_entry__.build() or make -C host)"
    scene = scenes.tiny_city_scene()
    scene_path = str(tmp_path / "scene.bin")
    scenes.save_scene_bin(scene, scene_path, W, H)
    out_path = str(tmp_path / "out.raw")
    cmd = [HOST, scene_path, renderer, str(W), str(H), str(FRAMES), out_path]
    if renderer == "nrc":
        params_path = str(tmp_path / "nrc.f16")
        probe = engine.NeuralRadianceCache(gfx_ctx, 2, 1e-2)
        num_params = probe.num_params
        probe.close()
        engine.random_nrc_params(num_params, 64 * 64 * 2 + 16 * 64, grid_amplitude=0.1).tofile(params_path)
        cmd.append(params_path)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    info = json.loads(res.stdout.strip().splitlines()[-1])
    assert info["renderer"] == renderer and info["kernel_launches"] > 0
    got = np.fromfile(out_path, dtype=np.float32).reshape(H, W, 4)
    want = _python_host(gfx_ctx, scene, renderer)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{renderer}: C++ host image differs from the Python host image"
    assert np.isfinite(got).all() and got[..., :3].mean() > 1e-4
