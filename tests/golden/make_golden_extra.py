#!/usr/bin/env python
"""CPU-only golden digests for the two "next" rows added at the end of round 1 (output side, instance animation):
python tests/golden/make_golden_extra.py writes tests/golden/golden_r01_extra.sha256 from the oracle.  Like golden_r01 these
freeze the checker, not the reference; tests/test_golden.py::test_oracle_reproduces_extra_golden replays them."""
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


def run_all() -> dict:
    out = {}
    scene = scenes.tiny_city_scene()
    osc = O.OracleScene(scene)
    fr = O.OracleFrame(osc, W, H)
    p = abi.default_frame_params(scene, W, H)
    p.log2NumCandidateSamples = 3
    current = list(scene.instances)
    for frame in range(3):
        if frame > 0:   # instances 1 and 2 move, rotate and grow every frame
            nxt = list(current)
            for k, i in enumerate((1, 2)):
                t = np.asarray(scene.instances[i].transform, dtype=np.float64)
                scale = float(np.linalg.norm(t[:, 0]))
                yaw = float(np.degrees(np.arctan2(t[0, 2], t[0, 0])))
                nxt[i] = scenes.move_instance(current[i], translate=(t[0, 3] + 0.2 * frame * (k + 1), t[1, 3], t[2, 3]),
                                              yaw_deg=yaw + 9.0 * frame, scale=scale * (1.0 + 0.04 * frame))
            current = nxt
            osc.update_instances(abi.make_instance_descs(current))
        for kind, pass_id in engine.restir_frame_passes(p, frame, 1, temporal=True, unbiased=False):
            fr.gbuffer(p) if kind == "gbuffer" else fr.restir(p, pass_id)
        out[f"animated/frame{frame}/beauty"] = fr.buffer(abi.BUF_BEAUTY_ACCUM, 0)
        out[f"animated/frame{frame}/motion"] = fr.buffer(abi.BUF_GBUFFER1, p.bufferIndex)
    beauty = fr.buffer(abi.BUF_BEAUTY_ACCUM, 0)
    out["present/beauty_tonemap_srgb"] = O.present(beauty, abi.PRESENT_COLOR, abi.PRESENT_TONE_MAP | abi.PRESENT_SRGB_GAMMA, 2.0, 1.0)
    out["present/beauty_flipped_raw_alpha"] = O.present(beauty, abi.PRESENT_COLOR, abi.PRESENT_FLIP_Y, 1.0, -1.0)
    out["present/normal"] = O.present(fr.buffer(abi.BUF_NORMAL_ACCUM, 0), abi.PRESENT_NORMAL, 0, 1.0, -1.0)
    return out


if __name__ == "__main__":
    res = run_all()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_r01_extra.sha256")
    with open(path, "w") as f:
        for name in sorted(res):
            f.write(f"{digest(res[name])}  {name}\n")
    print(f"wrote {len(res)} digests to {path}")
