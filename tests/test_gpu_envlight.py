"""GPU parity of the environment light's building blocks: the importance map built on upload
(RegularConstantContinuousDistribution2D, common_host.cpp:292-357) sampled and evaluated on the device, and the software
tex2DLod of the map, against the CPU oracle, bit for bit.  The renderers' use of it is covered by the *_env scenes of
test_gpu_restir / _restir_rearch / _pathtrace / _nrc_frame / _regir."""
import numpy as np
import pytest

from gfxexp_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def _device_query(ctx, op, pairs):
    import torch
    a = torch.from_numpy(np.ascontiguousarray(pairs, dtype=np.float32)).cuda()
    out = torch.zeros((a.shape[0], 3), dtype=torch.float32, device="cuda")
    ctx._check(ctx.lib.gfx_env_light_debug(ctx.h, None, op, a.data_ptr(), a.shape[0], out.data_ptr()), "gfx_env_light_debug")
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("size", [(64, 32), (37, 19)])  # a power of two and an odd size (the CDF search skips indices >= N)
def test_importance_map_and_fetch_bit_exact(gfx_ctx, oracle, size):
    scene = scenes.tiny_city_scene()
    scene.env_map = scenes.procedural_sky(*size)
    gfx_ctx.upload_scene(scene)
    oscene = oracle.OracleScene(scene, build_bvh=False)
    rng = np.random.default_rng(3)
    n = 100000
    u = rng.random((n, 2), dtype=np.float32)
    u[:64] = np.array([[0.0, 0.0], [0.99999994, 0.99999994], [0.5, 0.0], [0.0, 0.5]] * 16, dtype=np.float32)
    for op, what in ((0, "sample"), (1, "evaluatePDF"), (2, "fetch")):
        got = _device_query(gfx_ctx, op, u)
        want = oscene.env_query(op, u)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{what}: {np.argwhere(got != want)[:4].tolist()}"
    # fetch outside [0, 1): clamp addressing
    wide = (rng.random((4096, 2), dtype=np.float32) * 3 - 1).astype(np.float32)
    assert np.array_equal(_device_query(gfx_ctx, 2, wide).view(np.uint32), oscene.env_query(2, wide).view(np.uint32))


def test_no_environment_map_is_refused(gfx_ctx):
    gfx_ctx.upload_scene(scenes.tiny_city_scene())
    import torch
    a = torch.zeros((4, 2), device="cuda")
    out = torch.zeros((4, 3), device="cuda")
    assert gfx_ctx.lib.gfx_env_light_debug(gfx_ctx.h, None, 0, a.data_ptr(), 4, out.data_ptr()) != 0
