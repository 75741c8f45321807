#!/usr/bin/env python
"""Pins oracle/nrc.cpp's input encoding against the reference's own tiny-cuda-nn kernels (oracle/_ref/libtcnn_ref.so, built by
oracle/ref_tcnn/Makefile from /root/reference/ext/tiny-cuda-nn where it lies).  Runs on a GPU box, in its own process (it
launches third-party kernels: a fault there must not take the test session's CUDA context with it); prints one JSON line and
exits 0 when every compared half is bit-identical.  Usage: python tests/tcnn_ref_check.py"""
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib as O
from gfxexp_b200 import engine

LEVELS, FEATURES, LOG2_HASHMAP, BASE_RES, PER_LEVEL_SCALE = 16, 2, 15, 16, 2.0   # network_interface.cu:100-107


def level_offsets():
    """GridEncodingTemplated's constructor (grid.h:885-922): entries per level, multiple of 8, capped at 2^log2_hashmap_size"""
    offsets = [0]
    for level in range(LEVELS):
        scale = np.float32(np.exp2(np.float32(level * math.log2(PER_LEVEL_SCALE))) * BASE_RES - 1.0)
        res = int(math.ceil(float(scale))) + 1
        n = min(res ** 3, 0xFFFFFFFF // 2)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << LOG2_HASHMAP)
        offsets.append(offsets[-1] + n)
    return np.asarray(offsets, dtype=np.uint32)


def main():
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libtcnn_ref.so")
    ref = C.CDLL(ref_path)
    vp = C.c_void_p
    ref.tcnn_ref_grid_forward.argtypes = [C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_float, vp, C.c_size_t, vp, vp]
    ref.tcnn_ref_oneblob_forward.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]

    net = O.OracleNrc(2, 1e-2)
    params = engine.random_nrc_params(net.num_params, net.num_matrix_weights, seed=31, grid_amplitude=0.5)
    net.set_params(params)
    offsets = level_offsets()
    num_grid_params = int(offsets[-1]) * FEATURES
    assert net.num_matrix_weights + num_grid_params == net.num_params, "level table differs from the oracle's"
    grid = np.ascontiguousarray(np.asarray(params).view(np.uint16)[net.num_matrix_weights:])

    rng = np.random.default_rng(17)
    n = 128 * 64
    q = rng.uniform(0.0, 1.0, size=(n, 14)).astype(np.float32)
    q[:16, :3] = 0.0                       # cell corners and the upper boundary of the grid
    q[16:32, :3] = 1.0
    q[32:64, :3] = (rng.integers(0, 17, size=(32, 3)) / 16.0).astype(np.float32)
    q[64:96, 3:8] = rng.choice([0.0, 0.25, 0.5, 0.75, 1.0], size=(32, 5)).astype(np.float32)
    want = net.encode(q).view(np.uint16).reshape(n, 64)

    positions = np.ascontiguousarray(q[:, :3])
    got_grid = np.zeros((n, LEVELS * FEATURES), dtype=np.uint16)
    rc = ref.tcnn_ref_grid_forward(n, LEVELS, offsets.ctypes.data, BASE_RES, C.c_float(math.log2(PER_LEVEL_SCALE)), grid.ctypes.data,
                                   num_grid_params, positions.ctypes.data, got_grid.ctypes.data)
    assert rc == 0, "tcnn_ref_grid_forward failed"
    blob_in = np.ascontiguousarray(q[:, 3:8])
    got_blob = np.zeros((n, 20), dtype=np.uint16)
    rc = ref.tcnn_ref_oneblob_forward(n, 2, 5, blob_in.ctypes.data, got_blob.ctypes.data)
    assert rc == 0, "tcnn_ref_oneblob_forward failed"

    def compare(got, exp):
        bad = got != exp
        diff = np.abs(got.view(np.float16).astype(np.float64) - exp.view(np.float16).astype(np.float64))
        return {"mismatches": int(bad.sum()), "of": int(bad.size), "max_abs_diff": float(diff.max()),
                "first": np.argwhere(bad)[:3].tolist()}
    result = {"grid": compare(got_grid, want[:, :32]), "oneblob": compare(got_blob, want[:, 32:52])}
    print(json.dumps(result))
    return 0 if result["grid"]["mismatches"] == 0 and result["oneblob"]["mismatches"] == 0 else 3


if __name__ == "__main__":
    sys.exit(main())
