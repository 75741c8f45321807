#!/usr/bin/env python
"""Generates tests/golden/golden_r01.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference cannot be built or run here, so these are not reference outputs: they freeze what the
oracle produced when GPU parity was established, so that (a) a later edit of the oracle that changes any result is
caught by tests/test_golden.py on CPU, and (b) the GPU tests can check the CUDA path against fixed files in addition to
the live oracle.  Inputs are the seeded procedural scene `tiny_city_scene()` at 64x40 and the per-pixel RNG seed of
restir_di_main.cpp:1309-1321."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from gfxexp_b200 import abi, engine, scenes
from tests import oracle_lib as O

W, H = 64, 40


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class OracleBackend:
    """the operations make_golden / test_golden need, on the oracle"""

    def __init__(self, scene):
        self.scene = scene
        self.oscene = O.OracleScene(scene)
        self.frame = O.OracleFrame(self.oscene, W, H)

    def reset(self):
        self.frame = O.OracleFrame(self.oscene, W, H)

    def light_dist(self, i):
        pass

    def gbuffer(self, p): self.frame.gbuffer(p)
    def restir(self, p, pass_id):
        if pass_id >= abi.RESTIR_PRESAMPLE_LIGHTS:
            self.frame.restir_rearch(p, pass_id)
        else:
            self.frame.restir(p, pass_id)
    def pathtrace(self, p, variant): self.frame.pathtrace(p, variant)
    def regir_build(self, p, f, t): self.frame.regir_build_cells(p, f, t)
    def regir_update(self, p, f): self.frame.regir_update_access(p, f)
    def nrc_preprocess(self, p, a, b, new): self.frame.nrc_preprocess(p, a, b, new)
    def buffer(self, buf, idx=0): return self.frame.buffer(buf, idx)
    def linear(self, buf, idx=0, params=None): return self.frame.linear_buffer(buf, idx, params=params)


def run_all(be) -> dict:
    """every pipeline on the tiny scene; returns {name: array}"""
    out = {}
    scene = be.scene
    # --- ReSTIR DI, original renderer, biased and unbiased, 3 frames
    for unbiased in (False, True):
        be.reset()
        p = abi.default_frame_params(scene, W, H)
        for f in range(3):
            p.numAccumFrames = f
            be.light_dist(f % 2)
            for kind, pid in engine.restir_frame_passes(p, f, 1, True, unbiased):
                be.gbuffer(p) if kind == "gbuffer" else be.restir(p, pid)
        tag = "restir_unbiased" if unbiased else "restir_biased"
        out[tag + "_beauty"] = be.buffer(abi.BUF_BEAUTY_ACCUM)
        out[tag + "_rng"] = be.buffer(abi.BUF_RNG)
        out[tag + "_reservoir"] = be.buffer(abi.BUF_RESERVOIR, p.currentReservoirIndex)
    out["gbuffer0"] = be.buffer(abi.BUF_GBUFFER0, 0)
    out["gbuffer2"] = be.buffer(abi.BUF_GBUFFER2, 0)
    # --- rearchitected ReSTIR, 3 frames
    be.reset()
    p = abi.default_frame_params(scene, W, H)
    for f in range(3):
        p.numAccumFrames = f
        be.light_dist(f % 2)
        for kind, pid in engine.restir_rearch_frame_passes(p, f, True, True, False):
            be.gbuffer(p) if kind == "gbuffer" else be.restir(p, pid)
    out["rearch_beauty"] = be.buffer(abi.BUF_BEAUTY_ACCUM)
    out["rearch_sample_visibility"] = be.buffer(abi.BUF_SAMPLE_VISIBILITY, p.bufferIndex)
    # --- path tracer, 3 samples
    be.reset()
    p = abi.default_frame_params(scene, W, H)
    be.light_dist(0)
    for f in range(3):
        p.numAccumFrames = f
        be.gbuffer(p)
        be.pathtrace(p, abi.PT_BASELINE)
    out["pathtrace_beauty"] = be.buffer(abi.BUF_BEAUTY_ACCUM)
    out["pathtrace_rng"] = be.buffer(abi.BUF_RNG)
    # --- ReGIR, 3 frames on a 4x2x4 grid
    be.reset()
    p = abi.default_frame_params(scene, W, H)
    p.regirGridDim = (abi.c_u32 * 3)(4, 2, 4)
    be.light_dist(0)
    for f in range(3):
        p.frameIndex, p.bufferIndex, p.numAccumFrames = f, f % 2, f
        be.gbuffer(p)
        be.regir_build(p, f, f > 0)
        be.pathtrace(p, abi.PT_REGIR)
        be.regir_update(p, f)
    out["regir_beauty"] = be.buffer(abi.BUF_BEAUTY_ACCUM)
    out["regir_slots"] = be.linear(abi.BUF_REGIR_SLOTS, p.bufferIndex, params=p)
    # --- NRC path tracer (network-independent part), 2 frames
    be.reset()
    p = abi.default_frame_params(scene, W, H)
    be.light_dist(0)
    for f in range(2):
        p.frameIndex, p.bufferIndex, p.numAccumFrames = f, f % 2, f
        be.gbuffer(p)
        be.nrc_preprocess(p, 12345 + f, 678 + 5 * f, f == 0)
        be.pathtrace(p, abi.PT_NRC)
    st = be.linear(abi.BUF_NRC_STATE)[:, 0]
    ntrain = int(st[p.bufferIndex])
    out["nrc_state"] = st[:8].copy()
    out["nrc_contribution"] = be.linear(abi.BUF_NRC_FRAME_CONTRIBUTION)
    out["nrc_terminal_info"] = be.linear(abi.BUF_NRC_TERMINAL_INFO)
    out["nrc_train_query"] = be.linear(abi.BUF_NRC_TRAIN_QUERY, 0)[:ntrain]
    out["nrc_train_target"] = be.linear(abi.BUF_NRC_TRAIN_TARGET, 0)[:ntrain]
    out["nrc_vertex_info"] = be.linear(abi.BUF_NRC_TRAIN_VERTEX_INFO)[:ntrain]
    return out


def main():
    scene = scenes.tiny_city_scene()
    res = run_all(OracleBackend(scene))
    path = os.path.join(ROOT, "tests", "golden", "golden_r01.npz")
    # the images travel as arrays (a failing test can show where they differ), everything else as a digest
    np.savez_compressed(path, **{k: v for k, v in res.items() if k.endswith("_beauty") or k == "nrc_state"})
    with open(os.path.join(ROOT, "tests", "golden", "golden_r01.sha256"), "w") as f:
        for k in sorted(res):
            f.write(f"{digest(res[k])}  {k}  {res[k].dtype}  {list(res[k].shape)}\n")
    print(f"wrote {path}: {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
