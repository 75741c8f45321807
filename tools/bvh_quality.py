#!/usr/bin/env python
"""How much of the frame time is BVH quality?  Renders config-2 frames with (a) the GPU LBVH and (b) the CPU oracle's
SBVH (bvh::buildGeometryBVH<8> restatement) imported through gfx_bvh_import, and prints ms/frame plus the mean
traversal statistics of the primary rays for both.  Diagnostic tool (loads the oracle: not part of the product path)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gfxexp_b200 import abi, engine, scenes
from tests import oracle_lib


def measure(ctx, scene, w, h, tag):
    r = engine.ReSTIRRenderer(ctx, scene, w, h)
    for _ in range(3):
        r.render_frame()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        r.render_frame()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    rays = oracle_lib.primary_rays(r.params, w, h)
    hits = ctx.trace(rays[::7].copy(), abi.TRACE_CLOSEST | abi.TRACE_STATS)
    ud = hits["instUserData"]
    info = ctx.bvh_info()
    return {"bvh": tag, "ms_per_frame": ms, "nodes_per_primary_ray": float((ud & 0xFFFF).mean()),
            "tris_per_primary_ray": float((ud >> 16).mean()), "numNodes": info.numNodes, "numPrimRefs": info.numPrimRefs}


def main():
    small = "--small" in sys.argv
    scene = scenes.small_city_scene() if small else scenes.bistro_class_scene()
    w, h = (640, 360) if small else (1920, 1080)
    ctx = engine.Context(0)
    ctx.upload_scene(scene)
    variants = [("gpu-lbvh (GFX_BVH_BUILD_FAST)", 0x100), ("gpu-ploc (GFX_BVH_BUILD_PLOC)", 0x200), ("gpu-sah (default)", 0)]
    if "--ploc-only" in sys.argv:
        variants = variants[1:]
    if "--sweep" in sys.argv:
        variants = [(f"ploc r={r} maxLeaf={ml}", (r << 16) | ml) for r in (16, 32, 64) for ml in (2, 3, 4, 6)] + \
                   [(f"lbvh maxLeaf={ml}", 0x100 | ml) for ml in (2, 3, 6)]
    for tag, flags in variants:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ctx.build_bvh(flags)
        torch.cuda.synchronize()
        ev[0].record()
        ctx.build_bvh(flags)
        ev[1].record()
        torch.cuda.synchronize()
        ctx.create_frame(w, h)
        res = measure(ctx, scene, w, h, tag)
        res["build_ms"] = ev[0].elapsed_time(ev[1])
        print(json.dumps(res))
    if "--no-sbvh" in sys.argv:
        return
    if "--sah-nosplit" in sys.argv:
        cfg = oracle_lib.reference_build_config()
        cfg.splittingBudget = 0.0
        t = time.time()
        osc0 = oracle_lib.OracleScene(scene, cfg)
        nodes, refs, tris = osc0.export_bvh()
        print(f"oracle SAH (no spatial splits) build {time.time() - t:.1f}s", file=sys.stderr)
        ctx.import_bvh(nodes, refs, tris)
        ctx.create_frame(w, h)
        print(json.dumps(measure(ctx, scene, w, h, "oracle-sah-nosplit")))
    t = time.time()
    osc = oracle_lib.OracleScene(scene)
    nodes, refs, tris = osc.export_bvh()
    print(f"oracle SBVH build {time.time() - t:.1f}s", file=sys.stderr)
    ctx.import_bvh(nodes, refs, tris)
    ctx.create_frame(w, h)
    print(json.dumps(measure(ctx, scene, w, h, "oracle-sbvh")))


if __name__ == "__main__":
    main()
