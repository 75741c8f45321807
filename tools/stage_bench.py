#!/usr/bin/env python
"""Per-stage timings of the non-headline rows of SURVEY.md §8 on the config-2 scene at 1920x1080 (CUDA events around
each stage, inputs resident): path tracer (P1), NRC frame (N1-N5), ReGIR frame (C1), SVGF passes (V1-V4).
One JSON line per stage; `--small` uses the small city at 640x360 for a quick functional run."""
import json
import time
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gfxexp_b200 import abi, engine, scenes


class Timer:
    def __init__(self):
        self.acc = {}

    def run(self, name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.acc.setdefault(name, []).append((e0, e1))
        return out

    def result(self, skip=0):
        torch.cuda.synchronize()
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v[skip:]])) for k, v in self.acc.items()}


def main():
    small = "--small" in sys.argv
    scene_name = sys.argv[sys.argv.index("--scene") + 1] if "--scene" in sys.argv else None   # e.g. zero_day_class_scene (config 3)
    scene = getattr(scenes, scene_name)() if scene_name else (scenes.small_city_scene() if small else scenes.bistro_class_scene())
    w, h = (640, 360) if small else (1920, 1080)
    frames, warm = 12, 4
    ctx = engine.Context(0)
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(w, h)
    ctx.build_light_distributions(0)
    p = abi.default_frame_params(scene, w, h)
    npx = w * h

    def kernel_breakdown(fn, reps=4):
        """per-kernel ms per frame (CUDA events inside the library) of `fn`, averaged over `reps` calls"""
        ctx.timing_enable(True)
        ctx.timing_read()
        for i in range(reps):
            fn(i)
        t = ctx.timing_read()
        ctx.timing_enable(False)
        return {k: round(ms / reps, 4) for k, (ms, n) in sorted(t.items(), key=lambda kv: -kv[1][0])}

    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None

    def want(name):
        return only is None or name in only

    def stage_pathtrace():
        nonlocal p
        # ---- P1 path tracer
        t = Timer()
        ctx.read_stats(reset=True)
        for f in range(frames):
            p.numAccumFrames = f
            t.run("gbuffer", lambda: ctx.gbuffer(p))
            if f == warm:
                torch.cuda.synchronize()
                ctx.read_stats(reset=True)
            t.run("pathtrace", lambda: ctx.pathtrace(p))
        r = t.result(warm)
        rays = ctx.read_stats()[0] / (frames - warm)

        def pt_once(i):
            p.numAccumFrames = frames + i
            ctx.pathtrace(p)
        print(json.dumps({"stage": "pathtrace_baseline_kernels", **kernel_breakdown(pt_once)}))
        print(json.dumps({"stage": "pathtrace_baseline", "ms": r["pathtrace"], "gbuffer_ms": r["gbuffer"], "rays_per_frame": rays,
                          "Mrays_per_s": rays / ((r["pathtrace"] + r["gbuffer"]) * 1e-3) / 1e6, "width": w, "height": h}))

    if want("pathtrace"):
        stage_pathtrace()

    def stage_regir():
        nonlocal p
        # ---- C1 ReGIR
        t = Timer()
        p = abi.default_frame_params(scene, w, h)
        for f in range(frames):
            p.frameIndex, p.bufferIndex, p.numAccumFrames = f, f % 2, f
            ctx.gbuffer(p)
            t.run("build_cells", lambda: ctx.regir_build_cells(p, f, f > 0))
            t.run("pathtrace_regir", lambda: ctx.pathtrace(p, abi.PT_REGIR))
            t.run("update_access", lambda: ctx.regir_update_access(p, f))
        r = t.result(warm)

        def regir_once(i):
            f = frames + i
            p.frameIndex, p.bufferIndex, p.numAccumFrames = f, f % 2, f
            ctx.gbuffer(p)
            ctx.regir_build_cells(p, f, True)
            ctx.pathtrace(p, abi.PT_REGIR)
            ctx.regir_update_access(p, f)
        print(json.dumps({"stage": "regir_kernels", **kernel_breakdown(regir_once)}))
        active = int(ctx.download_linear(abi.BUF_REGIR_NUM_ACTIVE_CELLS, params=p)[(frames + 3) % 2, 0])
        slots = active * abi.REGIR_SLOTS_PER_CELL
        print(json.dumps({"stage": "regir", **r, "active_cells": active,
                          "build_GBps_algorithmic": 128.0 * slots / (r["build_cells"] * 1e-3) / 1e9}))

    if want("regir"):
        stage_regir()

    def stage_nrc():
        nonlocal p
        # ---- N1-N5 NRC frame
        t = Timer()
        p = abi.default_frame_params(scene, w, h)
        net = engine.NeuralRadianceCache(ctx, 2, 1e-2)
        net.set_params(engine.random_nrc_params(net.num_params, 64 * 64 * 2 + 16 * 64, grid_amplitude=0.1))
        rng = np.random.default_rng(0)
        nrc_frames = 24 if not small else frames
        for f in range(nrc_frames):
            p.frameIndex, p.bufferIndex, p.numAccumFrames = f, f % 2, f
            off = [int(rng.integers(0, 2 ** 32)) for _ in range(2)]
            ctx.gbuffer(p)
            t.run("preprocess", lambda: ctx.nrc_preprocess(p, off[0], off[1], f == 0))
            t.run("pathtrace_nrc", lambda: ctx.pathtrace(p, abi.PT_NRC))
            t.run("infer", lambda: ctx.nrc_frame_infer(net))
            t.run("accumulate", lambda: ctx.nrc_accumulate(p))
            t.run("propagate", lambda: ctx.nrc_propagate(p))
            t.run("shuffle", lambda: ctx.nrc_shuffle(p))
            t.run("train", lambda: ctx.nrc_frame_train(net))
        r = t.result(nrc_frames // 2)

        def nrc_once(i):
            f = nrc_frames + i
            p.numAccumFrames = f
            ctx.nrc_frame(net, p, f, [int(rng.integers(0, 2 ** 32)) for _ in range(2)], train=True)
        print(json.dumps({"stage": "nrc_kernels", **kernel_breakdown(nrc_once)}))
        st = ctx.download_linear(abi.BUF_NRC_STATE)[:, 0]
        b = (nrc_frames + 3) % 2
        nq = int(st[abi.NRC_STATE_NUM_INFERENCE_QUERIES])
        print(json.dumps({"stage": "nrc_frame", **r, "total_ms": sum(r.values()), "num_training_data": int(st[b]),
                          "tile": [int(st[2 + 2 * b]), int(st[3 + 2 * b])], "inference_queries": nq,
                          "infer_TFLOPs": 18432.0 * nq / (r["infer"] * 1e-3) / 1e12,
                          "train_TFLOPs": 55296.0 * 65536 / (r["train"] * 1e-3) / 1e12}))
        net.close()

    if want("nrc"):
        stage_nrc()

    def stage_svgf():
        nonlocal p
        # ---- V1-V4 SVGF on the ReSTIR frame (config 4)
        t = Timer()
        ren = engine.ReSTIRRenderer(ctx, scene, w, h)
        for f in range(frames):
            ren.render_frame()
            for pass_id, stage in engine.svgf_frame_passes(ren.params, f):
                name = {abi.SVGF_TEMPORAL_ACCUMULATE: "temporal", abi.SVGF_ESTIMATE_VARIANCE: "variance", abi.SVGF_ATROUS: "atrous",
                        abi.SVGF_FILL_BACKGROUND: "background", abi.SVGF_MODULATE_TAA: "modulate_taa"}[pass_id]
                t.run(name, lambda: ctx.svgf(ren.params, pass_id, stage))
        r = t.result(0)
        # a-trous runs 5x per frame: result() averaged per launch
        total = r["temporal"] + r["variance"] + 5 * r["atrous"] + r["background"] + r["modulate_taa"]
        print(json.dumps({"stage": "svgf", **{k + "_ms_per_launch": v for k, v in r.items()}, "total_ms_per_frame": total,
                          "GBps_algorithmic": 456.0 * npx / (total * 1e-3) / 1e9}))

    if want("svgf"):
        stage_svgf()

    def stage_combined():
        # ---- north-star target: ReSTIR DI (direct light at the primary hit) + NRC (indirect light) in one frame,
        # sharing the G-buffer: >= 60 fps at 1920x1080 on one GPU
        pc = abi.default_frame_params(scene, w, h)
        net = engine.NeuralRadianceCache(ctx, 2, 1e-2)
        net.set_params(engine.random_nrc_params(net.num_params, 64 * 64 * 2 + 16 * 64, grid_amplitude=0.1))
        rng = np.random.default_rng(0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n_frames, n_warm = 20, 6
        for f in range(n_frames + n_warm):
            if f == n_warm:
                ev[0].record()
            pc.numAccumFrames = f
            ctx.build_light_distributions(f % 2)
            for kind, pid in engine.restir_frame_passes(pc, f, 1, True, False):
                ctx.gbuffer(pc) if kind == "gbuffer" else ctx.restir(pc, pid)
            off = [int(rng.integers(0, 2 ** 32)) for _ in range(2)]
            ctx.nrc_preprocess(pc, off[0], off[1], f == 0)
            ctx.pathtrace(pc, abi.PT_NRC)
            ctx.nrc_frame_infer(net)
            ctx.nrc_accumulate(pc)
            ctx.nrc_propagate(pc)
            ctx.nrc_shuffle(pc)
            ctx.nrc_frame_train(net)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n_frames
        print(json.dumps({"stage": "restir_di_plus_nrc", "ms_per_frame": ms, "fps": 1e3 / ms, "width": w, "height": h,
                          "note": "one G-buffer, ReSTIR DI passes (32 candidates, temporal + 1x4 spatial), NRC path tracing + "
                                  "inference + 4 training steps"}))
        net.close()

    def stage_rearch():
        # ---- rearchitected ReSTIR DI (R5): presampled lights, per-tile subsets, <= 3 / 7 shadow rays per pixel
        for unbiased in (False, True):
            pr = abi.default_frame_params(scene, w, h)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            n_frames, n_warm = 20, 5
            for f in range(n_frames + n_warm):
                if f == n_warm:
                    ctx.timing_enable(True)
                    ctx.timing_read()
                    ev[0].record()
                pr.numAccumFrames = f
                ctx.build_light_distributions(f % 2)
                for kind, pid in engine.restir_rearch_frame_passes(pr, f, True, True, unbiased):
                    ctx.gbuffer(pr) if kind == "gbuffer" else ctx.restir(pr, pid)
            ev[1].record()
            torch.cuda.synchronize()
            per = {k: round(v[0] / n_frames, 4) for k, v in ctx.timing_read().items()}
            ctx.timing_enable(False)
            ms = ev[0].elapsed_time(ev[1]) / n_frames
            print(json.dumps({"stage": "restir_rearchitected_unbiased" if unbiased else "restir_rearchitected_biased",
                              "ms_per_frame_with_timing_events": ms, "fps": 1e3 / ms, "per_kernel_ms": per}))

    def stage_animated():
        # ---- every frame 10 % of the instances move (InstanceController::update), the BVH is rebuilt with the
        # LBVH builder and the light distributions are recomputed before the ReSTIR DI frame
        from gfxexp_b200 import scenes as sc
        pa = abi.default_frame_params(scene, w, h)
        movers = list(range(1, len(scene.instances), 10))
        current = list(scene.instances)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n_frames, n_warm = 10, 3
        descs_per_frame = []
        for f in range(n_frames + n_warm):
            nxt = list(current)
            for i in movers:
                t = np.asarray(current[i].transform, dtype=np.float64)
                scale = float(np.linalg.norm(t[:, 0]))
                yaw = float(np.degrees(np.arctan2(t[0, 2], t[0, 0])))
                nxt[i] = sc.move_instance(current[i], translate=(t[0, 3] + 0.02, t[1, 3], t[2, 3]), yaw_deg=yaw + 1.0, scale=scale)
            current = nxt
            descs_per_frame.append(abi.make_instance_descs(current))
        ctx.timing_enable(False)
        build_ms = []
        for f in range(n_frames + n_warm):
            if f == n_warm:
                torch.cuda.synchronize()
                ev[0].record()
            ctx.update_instances(descs_per_frame[f])
            t0 = time.perf_counter()
            ctx.build_bvh(0x100)
            if f >= n_warm:
                torch.cuda.synchronize()
                build_ms.append((time.perf_counter() - t0) * 1e3)
            ctx.build_light_distributions(f % 2)
            for kind, pid in engine.restir_frame_passes(pa, f, 1, True, False):
                ctx.gbuffer(pa) if kind == "gbuffer" else ctx.restir(pa, pid)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n_frames
        print(json.dumps({"stage": "restir_di_animated", "ms_per_frame": ms, "fps": 1e3 / ms, "moving_instances": len(movers),
                          "lbvh_rebuild_ms_wall": float(np.mean(build_ms)), "triangles": int(scene.num_triangles),
                          "note": "instance update + LBVH rebuild (GFX_BVH_BUILD_FAST) + light distributions + ReSTIR DI frame"}))
        ctx.upload_scene(scene)
        ctx.build_bvh()
        ctx.create_frame(w, h)

    if want("animated"):
        stage_animated()
    if want("rearch"):
        stage_rearch()
    if want("combined"):
        stage_combined()


if __name__ == "__main__":
    main()
