#!/usr/bin/env python
"""CPU-only golden digests for the rows added in round 2 (environment light, image textures, OBJ / MTL / PNG ingestion, the NRC
frame with an environment): python tests/golden/make_golden_r02.py writes tests/golden/golden_r02.sha256 from the oracle.  Like
golden_r01 these freeze the checker at the state in which the GPU parity tests were green, not the reference;
tests/test_golden.py::test_oracle_reproduces_r02_golden replays them."""
import hashlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from gfxexp_b200 import abi, engine, ingest, scenes
from tests import oracle_lib as O

W, H = 64, 40


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _restir(scene, out, tag, unbiased=False, frames=2):
    fr = O.OracleFrame(O.OracleScene(scene), W, H)
    p = abi.default_frame_params(scene, W, H)
    p.log2NumCandidateSamples = 3
    p.envLightRotation = 0.4
    for f in range(frames):
        p.numAccumFrames = f
        for kind, pid in engine.restir_frame_passes(p, f, 1, temporal=True, unbiased=unbiased):
            fr.gbuffer(p) if kind == "gbuffer" else fr.restir(p, pid)
    out[f"{tag}/beauty"] = fr.buffer(abi.BUF_BEAUTY_ACCUM)
    out[f"{tag}/albedo"] = fr.buffer(abi.BUF_ALBEDO_ACCUM)
    out[f"{tag}/reservoir"] = fr.buffer(abi.BUF_RESERVOIR, p.currentReservoirIndex)
    return fr, p


def run_all() -> dict:
    out = {}
    env = scenes.tiny_city_scene()
    env.env_map = scenes.procedural_sky(32, 16)
    _restir(env, out, "env/restir_biased")
    _restir(env, out, "env/restir_unbiased", unbiased=True)
    _restir(scenes.env_only_scene(), out, "env_only/restir")
    # path tracer and ReGIR under the environment
    fr = O.OracleFrame(O.OracleScene(env), W, H)
    p = abi.default_frame_params(env, W, H)
    for f in range(2):
        p.numAccumFrames, p.frameIndex, p.bufferIndex = f, f, f % 2
        fr.gbuffer(p)
        fr.pathtrace(p, abi.PT_BASELINE)
    out["env/pathtrace/beauty"] = fr.buffer(abi.BUF_BEAUTY_ACCUM)
    fr = O.OracleFrame(O.OracleScene(env), W, H)
    p = abi.default_frame_params(env, W, H)
    p.regirGridDim = (abi.c_u32 * 3)(6, 3, 6)
    for f in range(2):
        p.numAccumFrames, p.frameIndex, p.bufferIndex = f, f, f % 2
        fr.gbuffer(p)
        fr.regir_build_cells(p, f, f > 0)
        fr.pathtrace(p, abi.PT_REGIR)
        fr.regir_update_access(p, f)
    out["env/regir/beauty"] = fr.buffer(abi.BUF_BEAUTY_ACCUM)
    # the environment's building blocks on fixed inputs
    osc = O.OracleScene(env, build_bvh=False)
    u = np.random.default_rng(9).random((4096, 2), dtype=np.float32)
    for op, name in ((0, "sample"), (1, "pdf"), (2, "fetch")):
        out[f"env/{name}"] = osc.env_query(op, u)
    # image textures
    tex = scenes.tiny_city_scene()
    big = scenes.small_city_scene_textured()
    tex.textures = big.textures
    rng = np.random.default_rng(4)
    mt = np.full((tex.materials.shape[0], 4), 0xFFFFFFFF, dtype=np.uint32)
    for m in range(mt.shape[0]):
        if not tex.materials[m]["hasEmittance"]:
            mt[m, 0] = rng.integers(0, len(tex.textures))
            if rng.random() < 0.5 and tex.materials[m]["bsdfType"] != scenes.BSDF_LAMBERT:
                mt[m, 1], mt[m, 2] = 1, 2
    tex.material_textures = mt
    for mesh in tex.meshes:
        mesh.texcoords = (mesh.texcoords * np.float32(2.3) - np.float32(0.7)).astype(np.float32)
    _restir(tex, out, "textured/restir")
    fr = O.OracleFrame(O.OracleScene(tex), W, H)
    p = abi.default_frame_params(tex, W, H)
    for f in range(2):
        p.numAccumFrames, p.frameIndex, p.bufferIndex = f, f, f % 2
        fr.gbuffer(p)
        fr.pathtrace(p, abi.PT_BASELINE)
    out["textured/pathtrace/beauty"] = fr.buffer(abi.BUF_BEAUTY_ACCUM)
    # ingestion: the OBJ + MTL + PNG fixture of tests/test_ingest.py
    from tests.test_ingest import _write_scene
    import pathlib
    with tempfile.TemporaryDirectory() as d:
        _write_scene(pathlib.Path(d))
        scene = ingest.load_obj_scene(os.path.join(d, "test.obj"), camera_position=(0.0, 2.0, 6.0),
                                      camera_orientation=(scenes.rot_y(180.0) @ scenes.rot_x(15.0)))
    for i, m in enumerate(scene.meshes):
        out[f"ingest/mesh{i}/positions"] = m.positions
        out[f"ingest/mesh{i}/normals"] = m.normals
        out[f"ingest/mesh{i}/tangents"] = m.tangents
        out[f"ingest/mesh{i}/texcoords"] = m.texcoords
        out[f"ingest/mesh{i}/triangles"] = m.triangles
    out["ingest/materials"] = scene.materials.view(np.uint8)
    out["ingest/texture0"] = scene.textures[0]
    _restir(scene, out, "ingest/restir")
    return out


if __name__ == "__main__":
    res = run_all()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_r02.sha256")
    with open(path, "w") as f:
        for name in sorted(res):
            f.write(f"{digest(res[name])}  {name}\n")
    print(f"wrote {len(res)} digests to {path}")
