#!/usr/bin/env python
"""NRC network micro-benchmark (config 3 sizes): inference batch = pad128(1920*1080 + tiles) ~ 2.2 M queries,
training 4 x 16 384 samples (neural_radiance_caching_main.cpp:2304-2316,2350-2365).  Prints one JSON line:
ms per inference launch, achieved TFLOP/s on the padded MLP (18 432 FLOP/query) and compulsory-HBM GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gfxexp_b200 import engine


def main():
    ctx = engine.Context(0)
    net = engine.NeuralRadianceCache(ctx, 2, 1e-2)
    net.set_params(engine.random_nrc_params(net.num_params, 64 * 64 * 2 + 16 * 64, grid_amplitude=0.1))
    n = ((1920 * 1080 + 1920 * 1080 // 16 + 127) // 128) * 128
    q = torch.rand((n, 14), device="cuda")
    out = torch.empty((n, 3), device="cuda")
    for _ in range(3):
        net.infer(q, out, n)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 20
    ev[0].record()
    for _ in range(reps):
        net.infer(q, out, n)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    nt = 16384
    tq, tt = torch.rand((nt, 14), device="cuda"), torch.rand((nt, 3), device="cuda")
    for _ in range(3):
        net.train(tq, tt, nt)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        net.train(tq, tt, nt)
    ev[1].record()
    torch.cuda.synchronize()
    tms = ev[0].elapsed_time(ev[1]) / reps
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    flops = 18432.0 * n
    print(json.dumps({
        "nrc_infer_queries": n, "nrc_infer_ms": ms, "nrc_infer_tflops": flops / (ms * 1e-3) / 1e12,
        "tensor_frac_of_measured_bf16_peak": flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
        "nrc_infer_hbm_GBps": 68.0 * n / (ms * 1e-3) / 1e9, "hbm_frac": 68.0 * n / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
        "nrc_infer_Mqueries_per_s": n / (ms * 1e-3) / 1e6,
        "nrc_train_step_ms_16384": tms, "nrc_train_tflops": 55296.0 * nt / (tms * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
