"""The flattened light pick (gfxexp_b200/csrc/lights.cu k_pick*, lighting.cuh pickLightTriangle) against its definition: the
three nested DiscreteDistribution1D::sample calls of sampleLight (restir_di/restir_di_shared.h:356-409,
common/common_shared.h:209-246; lighting.cuh chainPickLightTriangle).  Both run on the GPU through gfx_light_pick_debug and
must return the same key for every float ul in [0, 1): the PCG32 lattice k * 2^-23, arbitrary floats (what a remapped ul, e.g.
after the environment-light split, looks like), denormals, bucket edges of the guide table and - the hard cases - the floats
around every piece boundary, which are found here by bisecting the chain on the host side of the comparison."""
import numpy as np
import pytest
import torch

from gfxexp_b200 import scenes

pytestmark = pytest.mark.gpu


def _pick(ctx, ul: np.ndarray):
    d_ul = torch.from_numpy(np.ascontiguousarray(ul, dtype=np.float32)).cuda()
    flat = torch.empty(d_ul.numel(), dtype=torch.int32, device="cuda")
    chain = torch.empty_like(flat)
    ctx._check(ctx.lib.gfx_light_pick_debug(ctx.h, None, d_ul.data_ptr(), d_ul.numel(), flat.data_ptr(), chain.data_ptr()),
               "gfx_light_pick_debug")
    torch.cuda.synchronize()
    return flat.cpu().numpy().view(np.uint32), chain.cpu().numpy().view(np.uint32)


def _check(ctx, ul, tag):
    flat, chain = _pick(ctx, ul)
    bad = np.flatnonzero(flat != chain)
    assert bad.size == 0, (f"{tag}: {bad.size} of {ul.size} picks differ, first ul bits "
                           f"{ul.view(np.uint32)[bad[:4]].tolist()} flat {flat[bad[:4]].tolist()} chain {chain[bad[:4]].tolist()}")
    return chain


@pytest.mark.parametrize("scene_name", ["small_city_scene", "bistro_class_scene"])
def test_flat_pick_equals_the_sampling_chain(gfx_ctx, scene_name):
    scene = getattr(scenes, scene_name)()
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_light_distributions()
    rng = np.random.default_rng(5)
    n = 1 << 21
    # the PCG32 lattice: (bits >> 9 | 0x3f800000) - 1
    lattice = (rng.integers(0, 1 << 23, size=n, dtype=np.uint32) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    keys = _check(gfx_ctx, lattice, "lattice")
    assert np.unique(keys).size > min(scene.num_emissive_triangles // 2, 1000), "the scene's lights are not reached"
    # one contiguous run of the lattice (every piece boundary inside it is crossed)
    run = (np.arange(5_000_000, 5_000_000 + n, dtype=np.uint32) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    _check(gfx_ctx, run, "lattice run")
    # arbitrary floats in [0, 1), uniformly distributed bit patterns (mostly tiny values and denormals), and the top of the range
    _check(gfx_ctx, rng.random(n, dtype=np.float32), "uniform floats")
    _check(gfx_ctx, rng.integers(0, 0x3F800000, size=n, dtype=np.uint32).view(np.float32), "uniform bit patterns")
    _check(gfx_ctx, np.arange(0x3F800000 - n, 0x3F800000, dtype=np.uint32).view(np.float32), "top of the range")
    _check(gfx_ctx, np.arange(0, n, dtype=np.uint32).view(np.float32), "denormals")
    # guide-bucket edges +- 2 floats
    j = np.arange(1, 1 << 17, dtype=np.uint32)
    edges = (j.astype(np.float32) / np.float32(1 << 17)).view(np.uint32).astype(np.int64)
    around = np.concatenate([edges + d for d in (-2, -1, 0, 1, 2)]).astype(np.uint32)
    _check(gfx_ctx, around.view(np.float32), "bucket edges")
    # piece boundaries: bisect between lattice neighbours with different keys down to adjacent floats, check +-2 around them
    order = np.argsort(lattice.view(np.uint32))
    bits = lattice.view(np.uint32)[order]
    k = keys[order]
    change = np.flatnonzero(k[1:] != k[:-1])
    lo, hi = bits[change].astype(np.int64), bits[change + 1].astype(np.int64)
    klo = k[change]
    for _ in range(32):
        mid = ((lo + hi) // 2).astype(np.uint32)
        _, kmid = _pick(gfx_ctx, mid.view(np.float32))
        same = kmid == klo
        lo = np.where(same, mid, lo).astype(np.int64)
        hi = np.where(same, hi, mid).astype(np.int64)
    assert np.all(hi - lo <= 1)
    near = np.concatenate([(hi + d) for d in (-3, -2, -1, 0, 1, 2)])
    near = np.clip(near, 0, 0x3F7FFFFF).astype(np.uint32)
    _check(gfx_ctx, near.view(np.float32), f"{change.size} piece boundaries")
