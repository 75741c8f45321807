#!/usr/bin/env python
"""What fraction of the RIS candidates is dark, and why - measured with the CPU oracle (no GPU needed).  The CUDA candidate
kernel is bound by L1 wavefronts of the light-triangle fetch (profiles/r01_summary.md); these fractions say how much a staged
fetch or a per-triangle cull record can save.  Usage: python tools/ris_candidate_stats.py [--config2]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O

from gfxexp_b200 import abi, engine, scenes


def main():
    big = "--config2" in sys.argv
    scene = scenes.bistro_class_scene() if big else scenes.small_city_scene()
    w, h = (1920, 1080) if big else (320, 200)
    osc = O.OracleScene(scene)
    fr = O.OracleFrame(osc, w, h)
    p = abi.default_frame_params(scene, w, h)
    if big:   # a 64-row strip through the middle of the frame is enough
        p.tileOriginY, p.tileRows = 500, 64
    lib = O.lib()
    lib.orc_ris_stats.argtypes = [C.c_int, C.c_void_p]
    lib.orc_ris_stats(1, None)
    for kind, pass_id in engine.restir_frame_passes(p, 0, 0):
        fr.gbuffer(p) if kind == "gbuffer" else fr.restir(p, pass_id)
    out = (C.c_ulonglong * 5)()
    lib.orc_ris_stats(0, out)
    n = sum(out)
    names = ["zero_density", "below_horizon", "faces_away", "otherwise_zero", "contributing"]
    print(json.dumps({"scene": "config 2 (64-row strip)" if big else "small_city 320x200", "candidates": int(n),
                      **{k: round(out[i] / max(n, 1), 4) for i, k in enumerate(names)}}))


if __name__ == "__main__":
    main()
