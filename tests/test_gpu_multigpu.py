"""2-GPU parity (skipped on a 1-GPU box): the strip-sharded frame composited on rank 0 is bit-identical to the single-GPU
frame of the same library, with the seam rows exchanged by one-sided pushes over NVLink peer memory (csrc/peer.cu) and
with NCCL send/recv."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, FRAMES = 320, 192, 3


def _worker(rank, world, port, result_path, peer):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gfxexp_b200 import abi, engine, multigpu, scenes
    scene = scenes.small_city_scene()
    ctx = engine.Context(rank)
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(W, H)
    p = abi.default_frame_params(scene, W, H)
    driver = multigpu.StripDriver(ctx, p, W, H, rank, world, peer=peer)
    assert driver.backend.peer_ready == peer
    outs = []
    for f in range(FRAMES):
        driver.render_frame(f, num_spatial_passes=2)
        torch.cuda.synchronize()
        outs.append(driver.composited.cpu().numpy().copy())
    assert not ctx.peer_timed_out()
    if rank == 0:
        np.save(result_path, np.stack(outs))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


@pytest.mark.parametrize("peer", [True, False])
def test_two_gpu_strips_equal_single_gpu(tmp_path, gfx_ctx, peer):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from gfxexp_b200 import abi, engine, scenes
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    result = str(tmp_path / "composited.npy")
    mp.spawn(_worker, args=(2, port, result, peer), nprocs=2, join=True)
    got = np.load(result)

    scene = scenes.small_city_scene()
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    gfx_ctx.create_frame(W, H)
    p = abi.default_frame_params(scene, W, H)
    for f in range(FRAMES):
        gfx_ctx.build_light_distributions(f % 2)
        for kind, pass_id in engine.restir_frame_passes(p, f, 2):
            if kind == "gbuffer":
                gfx_ctx.gbuffer(p)
            else:
                gfx_ctx.restir(p, pass_id)
        gfx_ctx.synchronize()
        want = gfx_ctx.download(abi.BUF_BEAUTY_ACCUM).reshape(-1)
        assert np.array_equal(got[f].view(np.uint32), want.view(np.uint32)), f"frame {f}: 2-GPU frame differs"
