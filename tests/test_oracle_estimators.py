"""Pins the oracle against itself statistically: five independently restated estimators of DIRECT lighting must agree in
expectation (the check the reference's restir_di/RIS_Test/ris_test.ipynb makes for RIS in 1-D, here on the real pipeline).
  (a) path tracer with maxPathLength 2: NEE + implicit hits, power-heuristic MIS    (path_tracing)
  (b) ReSTIR initial RIS (32 candidates) + shading, no reuse                        (restir_di, original)
  (c) ReSTIR unbiased: temporal + spatial reuse with the unbiased contribution MIS  (restir_di, original)
  (d) rearchitected ReSTIR unbiased: presampled lights, pairwise MIS                (restir_di, rearchitected)
  (e) ReGIR with maxPathLength 2: reservoir-grid NEE                                (regir)
A wrong pdf, MIS weight, M cap or reservoir weight in any of them shows up as a few-percent difference of the image mean;
the run-to-run noise of the mean at 384 accumulated frames is ~0.5 %."""
import numpy as np

from gfxexp_b200 import abi, engine, scenes

W, H, N = 48, 30, 384


def _render(oracle, oscene, scene, kind):
    fr = oracle.OracleFrame(oscene, W, H)
    p = abi.default_frame_params(scene, W, H)
    if kind in ("pt2", "regir2"):
        p.maxPathLength = 2
    if kind == "regir2":
        p.regirGridDim = (abi.c_u32 * 3)(8, 4, 8)
    if kind == "ris":
        p.enableTemporalReuse = 0
        p.enableSpatialReuse = 0
    for f in range(N):
        p.numAccumFrames = f
        if kind == "pt2":
            fr.gbuffer(p)
            fr.pathtrace(p, abi.PT_BASELINE)
        elif kind == "ris":
            p.frameIndex, p.bufferIndex, p.currentReservoirIndex = f, f % 2, 0
            fr.gbuffer(p)
            fr.restir(p, abi.RESTIR_INITIAL_RIS)
            fr.restir(p, abi.RESTIR_SHADING)
        elif kind == "restir_unbiased":
            for k, pid in engine.restir_frame_passes(p, f, 1, True, True):
                fr.gbuffer(p) if k == "gbuffer" else fr.restir(p, pid)
        elif kind == "rearch_unbiased":
            for k, pid in engine.restir_rearch_frame_passes(p, f, True, True, True):
                fr.gbuffer(p) if k == "gbuffer" else fr.restir_rearch(p, pid)
        elif kind == "regir2":
            p.frameIndex, p.bufferIndex = f, f % 2
            fr.gbuffer(p)
            fr.regir_build_cells(p, f, f > 0)
            fr.pathtrace(p, abi.PT_REGIR)
            fr.regir_update_access(p, f)
    hit = fr.buffer(abi.BUF_GBUFFER0, 0)[..., 0] != 0xFFFFFFFF  # misses get different background constants (0.001 / 0.01)
    return fr.buffer(abi.BUF_BEAUTY_ACCUM)[..., :3][hit]


def test_direct_lighting_estimators_agree(oracle):
    scene = scenes.tiny_city_scene()
    oscene = oracle.OracleScene(scene)
    images = {k: _render(oracle, oscene, scene, k) for k in ("pt2", "ris", "restir_unbiased", "rearch_unbiased", "regir2")}
    ref = float(images["pt2"].mean())
    assert ref > 1e-3
    for k, img in images.items():
        assert np.isfinite(img).all()
        rel = (float(img.mean()) - ref) / ref
        assert abs(rel) < 0.025, f"{k}: image mean differs from the path tracer's by {100 * rel:.2f} %"
