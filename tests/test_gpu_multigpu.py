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


def _worker_allgather(rank, world, port, result_path):
    """gfx_framebuffer_allgather with the HOST's own ncclComm_t (created here through ctypes on the NCCL torch ships, the way a
    C++ host would with nccl.h): every rank fills its strip of the beauty buffer, the library all-gathers in place"""
    import ctypes as C
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gfxexp_b200 import abi, engine, scenes
    nccl = C.CDLL("libnccl.so.2")  # the soname resolves to the copy torch has loaded

    class UniqueId(C.Structure):  # ncclUniqueId, passed by value
        _fields_ = [("internal", C.c_byte * 128)]
    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    ids = [bytes(uid)]
    dist.broadcast_object_list(ids, src=0)
    C.memmove(C.byref(uid), ids[0], 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    ctx = engine.Context(rank)
    ctx.upload_scene(scenes.tiny_city_scene())
    ctx.create_frame(W, H)
    rows = H // world
    beauty = np.zeros((H, W, 4), dtype=np.float32)
    beauty[rank * rows:(rank + 1) * rows] = rank + 1.0
    ctx.upload(abi.BUF_BEAUTY_ACCUM, 0, beauty)
    ptr, _ = ctx.device_ptr(abi.BUF_BEAUTY_ACCUM, 0)
    ctx._check(ctx.lib.gfx_framebuffer_allgather(ctx.h, comm, None, rows, ptr), "gfx_framebuffer_allgather")
    torch.cuda.synchronize()
    got = ctx.download(abi.BUF_BEAUTY_ACCUM)
    np.save(result_path + f".{rank}.npy", got)
    dist.barrier()
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.destroy_process_group()
    ctx.close()


def test_framebuffer_allgather_through_the_c_abi(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    result = str(tmp_path / "gathered")
    mp.spawn(_worker_allgather, args=(2, port, result), nprocs=2, join=True)
    rows = H // 2
    for rank in range(2):
        got = np.load(result + f".{rank}.npy")
        assert np.all(got[:rows] == 1.0) and np.all(got[rows:] == 2.0), f"rank {rank}: gathered frame is wrong"


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
