#!/usr/bin/env python
"""Cost profile of the config-2 frame by screen rows, measured on ONE GPU: every strip of an N-way split is rendered alone
(tileOriginY / tileRows, no seam exchange: wrong pixels near the seams, same work) after a few full warm-up frames.
Prints the per-strip milliseconds, so that the strong-scaling ceiling max(strip) and the benefit of cost-balanced strip
boundaries can be read off without a multi-GPU box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gfxexp_b200 import abi, engine, scenes

W, H = 1920, 1080


def frame(ctx, p, fi, lo, hi):
    ctx.build_light_distributions(fi % 2)
    for kind, pass_id in engine.restir_frame_passes(p, fi, 1):
        p.tileOriginY, p.tileRows = lo, hi - lo
        if kind == "gbuffer":
            ctx.gbuffer(p)
        else:
            ctx.restir(p, pass_id)
    p.tileOriginY, p.tileRows = 0, 0


def main():
    scene = scenes.bistro_class_scene()
    ctx = engine.Context(0)
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(W, H)
    p = abi.default_frame_params(scene, W, H)
    fi = 0
    for _ in range(4):
        frame(ctx, p, fi, 0, H)
        fi += 1
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(lo, hi, reps=6):
        nonlocal fi
        frame(ctx, p, fi, lo, hi); fi += 1
        ev[0].record()
        for _ in range(reps):
            frame(ctx, p, fi, lo, hi)
            fi += 1
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    if "--fixed" in sys.argv:
        # fixed per-frame cost: per-kernel event times of an 8-row and a 135-row strip
        for rows in (8, 135, H):
            timed(400, 400 + rows if rows < H else H, 2) if rows < H else timed(0, H, 2)
            ctx.timing_enable(True)
            ctx.timing_read()
            reps = 5
            lo, hi = (400, 400 + rows) if rows < H else (0, H)
            t = timed(lo, hi, reps)
            per = {k: round(v[0] / (reps + 1), 4) for k, v in ctx.timing_read().items()}
            ctx.timing_enable(False)
            print(json.dumps({"rows": rows, "ms_frame_with_timing_events": t, "kernel_sum": round(sum(per.values()), 4), "per_kernel": per}))
        return
    full = timed(0, H)
    print(json.dumps({"strips": 1, "ms": [full]}))
    rows120 = [timed(y, y + 120) for y in range(0, H, 120)]
    print(json.dumps({"rows_of_120": rows120, "sum": sum(rows120)}))
    for n in (2, 4, 8):
        r = H // n
        ms = [timed(k * r, (k + 1) * r) for k in range(n)]
        print(json.dumps({"strips": n, "ms": ms, "max": max(ms), "ideal": full / n, "efficiency_ceiling": full / n / max(ms)}))


if __name__ == "__main__":
    main()
