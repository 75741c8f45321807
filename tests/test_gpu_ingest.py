"""GPU parity on an ingested scene: OBJ + MTL + PNG through gfxexp_b200.ingest, rendered by the library
and by the oracle - G-buffers (albedo through the floor's image texture), ReSTIR DI and the path tracer, bit for bit."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, ingest, scenes
from tests.test_ingest import _write_scene

pytestmark = pytest.mark.gpu


def _same(got, want, tag):
    g = got.view(np.uint32) if got.dtype != np.uint64 else got
    w = want.view(np.uint32) if want.dtype != np.uint64 else want
    assert np.array_equal(g, w), f"{tag}: {np.argwhere(g != w)[:4].tolist()}"


def test_ingested_obj_scene_bit_exact(gfx_ctx, oracle, tmp_path):
    _write_scene(tmp_path)
    scene = ingest.load_obj_scene(str(tmp_path / "test.obj"), camera_position=(0.0, 2.0, 6.0),
                                  camera_orientation=(scenes.rot_y(180.0) @ scenes.rot_x(15.0)))
    w, h = 96, 64
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    gfx_ctx.create_frame(w, h)
    oframe = oracle.OracleFrame(oracle.OracleScene(scene), w, h)
    p = abi.default_frame_params(scene, w, h)
    for f in range(3):
        p.numAccumFrames = f
        gfx_ctx.build_light_distributions(f % 2)
        for kind, pid in engine.restir_frame_passes(p, f, 1):
            if kind == "gbuffer":
                gfx_ctx.gbuffer(p)
                oframe.gbuffer(p)
            else:
                gfx_ctx.restir(p, pid)
                oframe.restir(p, pid)
        for buf in (abi.BUF_GBUFFER3, abi.BUF_ALBEDO_ACCUM, abi.BUF_RNG, abi.BUF_BEAUTY_ACCUM):
            _same(gfx_ctx.download(buf, p.bufferIndex if buf == abi.BUF_GBUFFER3 else 0),
                  oframe.buffer(buf, p.bufferIndex if buf == abi.BUF_GBUFFER3 else 0), f"restir frame {f} buffer {buf}")
    for f in range(2):
        p.numAccumFrames, p.frameIndex, p.bufferIndex = f, f, f % 2
        gfx_ctx.gbuffer(p)
        gfx_ctx.pathtrace(p)
        oframe.gbuffer(p)
        oframe.pathtrace(p)
        _same(gfx_ctx.download(abi.BUF_BEAUTY_ACCUM), oframe.buffer(abi.BUF_BEAUTY_ACCUM), f"path tracer frame {f}")
    assert gfx_ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3].mean() > 1e-3


def test_textured_emitter_is_refused(gfx_ctx):
    scene = scenes.small_city_scene_textured()
    emitter = int(np.argmax(scene.materials["hasEmittance"]))
    scene.material_textures[emitter, 3] = 0
    with pytest.raises(engine.GfxError):
        gfx_ctx.upload_scene(scene)
