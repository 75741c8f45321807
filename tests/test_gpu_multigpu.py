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


def _worker(rank, world, port, result_path, peer, balanced=False):
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
    boundaries = None
    if balanced:  # strips of equal work instead of equal height, composited by gfx_framebuffer_allgatherv on a raw communicator
        boundaries = multigpu.StripDriver.cost_balanced_boundaries(ctx, p, W, H, world)
        if boundaries[1] == H // 2:  # the view happens to be balanced by equal strips: still exercise unequal ones
            boundaries = [0, H // 2 + 8, H]
    driver = multigpu.StripDriver(ctx, p, W, H, rank, world, peer=peer, boundaries=boundaries)
    if balanced:
        driver.use_raw_communicator()
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


NRC_FRAMES = 4


def _nrc_offsets(f):
    rng = np.random.default_rng(1000 + f)
    return [int(rng.integers(0, 2 ** 32)) for _ in range(2)]


def _worker_restir_nrc(rank, world, port, result_path, boundaries=None):
    """config 5 in small: ReSTIR DI + NRC in one frame, sharded by strips; the NRC half numbers its training vertices over
    the whole frame (one word per rank and round all-gathered), merges the records and trains replicated"""
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
    p.maxPathLength = 5
    net = engine.NeuralRadianceCache(ctx, 2, 1e-2)
    driver = multigpu.StripDriver(ctx, p, W, H, rank, world, boundaries=boundaries)
    driver.enable_nrc(net)
    outs, states = [], []
    for f in range(NRC_FRAMES):
        p.numAccumFrames = f
        driver.render_restir_nrc_frame(f, _nrc_offsets(f))
        torch.cuda.synchronize()
        outs.append(driver.composited.cpu().numpy().copy())
        states.append(ctx.download_linear(abi.BUF_NRC_STATE)[:, 0].copy())
    assert not ctx.peer_timed_out()
    np.save(result_path + f".weights.{rank}.npy", net.read(abi.NRC_READ_MASTER))
    np.save(result_path + f".state.{rank}.npy", np.stack(states))
    if rank == 0:
        np.save(result_path + ".frames.npy", np.stack(outs))
    dist.barrier()
    net.close()
    driver.comm.close()
    dist.destroy_process_group()
    ctx.close()


@pytest.mark.parametrize("boundaries", [None, [0, 72, 192]])
def test_restir_nrc_strips_equal_single_gpu(tmp_path, gfx_ctx, boundaries):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from gfxexp_b200 import abi, engine, scenes
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    result = str(tmp_path / "restir_nrc")
    mp.spawn(_worker_restir_nrc, args=(2, port, result, boundaries), nprocs=2, join=True)
    got = np.load(result + ".frames.npy")

    scene = scenes.small_city_scene()
    ctx = gfx_ctx
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(W, H)
    p = abi.default_frame_params(scene, W, H)
    p.maxPathLength = 5
    net = engine.NeuralRadianceCache(ctx, 2, 1e-2)
    states = []
    for f in range(NRC_FRAMES):
        p.numAccumFrames = f
        ctx.build_light_distributions(f % 2)
        for kind, pass_id in engine.restir_frame_passes(p, f, 1):
            ctx.gbuffer(p) if kind == "gbuffer" else ctx.restir(p, pass_id)
        off = _nrc_offsets(f)
        ctx.nrc_preprocess(p, off[0], off[1], f == 0)
        ctx.pathtrace(p, abi.PT_NRC)
        ctx.nrc_frame_infer(net)
        ctx.nrc_accumulate(p)
        ctx.nrc_propagate(p)
        ctx.nrc_shuffle(p)
        ctx.nrc_frame_train(net)
        ctx.synchronize()
        want = ctx.download(abi.BUF_BEAUTY_ACCUM).reshape(-1)
        states.append(ctx.download_linear(abi.BUF_NRC_STATE)[:, 0].copy())
        assert np.array_equal(got[f].view(np.uint32), want.view(np.uint32)), f"frame {f}: sharded DI + NRC frame differs"
    want_weights = net.read(abi.NRC_READ_MASTER)
    net.close()
    states = np.stack(states)
    assert states[-1][abi.NRC_STATE_NUM_TRAINING_DATA + (NRC_FRAMES - 1) % 2] > 0, "no training data: the test is vacuous"
    for rank in range(2):
        w = np.load(result + f".weights.{rank}.npy")
        assert np.array_equal(w.view(np.uint32), want_weights.view(np.uint32)), f"rank {rank}: trained weights differ"
        st = np.load(result + f".state.{rank}.npy")
        # training-data counters and tile sizes of both buffers: words 0..5
        assert np.array_equal(st[:, :6], states[:, :6]), f"rank {rank}: training-vertex counters / tile sizes differ"


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


@pytest.mark.parametrize("peer,balanced", [(True, False), (False, False), (True, True)])
def test_two_gpu_strips_equal_single_gpu(tmp_path, gfx_ctx, peer, balanced):
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
    mp.spawn(_worker, args=(2, port, result, peer, balanced), nprocs=2, join=True)
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
