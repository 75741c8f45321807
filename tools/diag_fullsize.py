#!/usr/bin/env python
"""Diagnostic: frame 0 of config 2 at full size, pass by pass, against the oracle on rows 500..579 (GPU box).
Prints the number of differing 32-bit words per buffer after each pass.  GFXB200_LIB selects an A/B build."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from gfxexp_b200 import abi, engine, scenes
from tests import oracle_lib as O

W, H = 1920, 1080
scene = scenes.bistro_class_scene()
ctx = engine.Context(0)
ctx.upload_scene(scene)
ctx.build_bvh(0x200)
nodes, refs, tris = ctx.export_bvh()
ctx.build_bvh(0)
checker = O.OracleScene(scene, build_bvh=False)
checker.import_bvh(nodes, refs, tris)
ctx.create_frame(W, H)
oframe = O.OracleFrame(checker, W, H)
p = abi.default_frame_params(scene, W, H)
po = abi.default_frame_params(scene, W, H)
po.tileOriginY, po.tileRows = 500, 80
ctx.build_light_distributions(0)
bufs = [(abi.BUF_GBUFFER0, 0), (abi.BUF_GBUFFER2, 0), (abi.BUF_GBUFFER3, 0), (abi.BUF_RNG, 0), (abi.BUF_RESERVOIR, 0), (abi.BUF_RESERVOIR, 1),
        (abi.BUF_RESERVOIR_INFO, 0), (abi.BUF_RESERVOIR_INFO, 1), (abi.BUF_BEAUTY_ACCUM, 0)]
gen_g = engine.restir_frame_passes(p, 0, 1, True, False)
gen_o = engine.restir_frame_passes(po, 0, 1, True, False)
for (kind, pid), (_, _) in zip(gen_g, gen_o):
    if kind == "gbuffer":
        ctx.gbuffer(p)
        oframe.gbuffer(po)
    else:
        ctx.restir(p, pid)
        oframe.restir(po, pid)
    ctx.synchronize()
    line = []
    for buf, idx in bufs:
        got, want = ctx.download(buf, idx), oframe.buffer(buf, idx)
        g = got[530:550] if got.shape[0] == H else got.reshape(-1, H, W, got.shape[-1])[:, 530:550]
        w = want[530:550] if want.shape[0] == H else want.reshape(-1, H, W, want.shape[-1])[:, 530:550]
        gv = g.view(np.uint32) if g.dtype != np.uint64 else g
        wv = w.view(np.uint32) if w.dtype != np.uint64 else w
        line.append(f"{buf}[{idx}]:{int((gv != wv).sum())}")
    print(f"after {kind} {pid}: " + " ".join(line), flush=True)
