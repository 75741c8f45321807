#!/usr/bin/env python
"""Path tracer micro-benchmark: config-2 scene (bistro-class city, 1920x1080), 1 spp, maxPathLength 5.
Prints one JSON line: ms per frame of gfx_pathtrace_launch (G-buffer excluded), rays per frame, Mrays/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gfxexp_b200 import abi, engine, scenes


def main():
    small = "--small" in sys.argv
    scene = scenes.small_city_scene() if small else scenes.bistro_class_scene()
    w, h = (640, 360) if small else (1920, 1080)
    ctx = engine.Context(0)
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(w, h)
    ctx.build_light_distributions(0)
    p = abi.default_frame_params(scene, w, h)
    ctx.gbuffer(p)
    for i in range(3):
        p.numAccumFrames = i
        ctx.pathtrace(p)
    torch.cuda.synchronize()
    ctx.read_stats(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 10
    ev[0].record()
    for i in range(reps):
        p.numAccumFrames = 3 + i
        ctx.pathtrace(p)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    rays = ctx.read_stats()[0] / reps
    print(json.dumps({"pathtrace_ms": ms, "rays_per_frame": rays, "rays_per_px": rays / (w * h),
                      "Mrays_per_s": rays / (ms * 1e-3) / 1e6, "width": w, "height": h}))


if __name__ == "__main__":
    main()
