"""Scene ingestion: OBJ + MTL + PNG -> Scene with the conventions the reference asks of assimp
(common/common_host.cpp:2178-2429), and the CPU oracle renders the result with its image textures."""
import math
import struct
import zlib

import numpy as np

from gfxexp_b200 import abi, engine, ingest, scenes
from tests import oracle_lib as O

OBJ = """# a textured floor quad (shared vertices, vt, vn), a pyramid without normals, an emitter quad above
mtllib test.mtl
v -2 0 -2
v  2 0 -2
v  2 0  2
v -2 0  2
vt 0 0
vt 2 0
vt 2 2
vt 0 2
vn 0 1 0
usemtl floor
f 1/1/1 4/4/1 3/3/1 2/2/1
v -0.5 0 -0.5
v  0.5 0 -0.5
v  0.5 0  0.5
v -0.5 0  0.5
v  0 1 0
usemtl stone
f 5 9 6
f 6 9 7
f 7 9 8
f 8 9 5
v -0.5 3 -0.5
v  0.5 3 -0.5
v  0.5 3  0.5
v -0.5 3  0.5
usemtl lamp
f -4 -3 -2 -1
"""
MTL = """newmtl floor
Kd 1 1 1
Ks 0.04 0.04 0.04
Ns 25
map_Kd floor.png
newmtl stone
Kd 0.5 0.4 0.3
Ks 0.2 0.2 0.2
Ns 100
newmtl lamp
Kd 0 0 0
Ke 30 28 25
"""


def _png_with_filters(img):
    """PNG writer that cycles through all five row filters (the decoder's inverse paths)"""
    h, w, c = img.shape
    raw = bytearray()
    prev = np.zeros(w * c, dtype=np.int32)
    for y in range(h):
        line = img[y].reshape(-1).astype(np.int32)
        f = y % 5
        out = np.zeros_like(line)
        for i in range(len(line)):
            a = line[i - c] if i >= c else 0
            b = prev[i]
            cc = prev[i - c] if i >= c else 0
            if f == 0:
                pred = 0
            elif f == 1:
                pred = a
            elif f == 2:
                pred = b
            elif f == 3:
                pred = (a + b) >> 1
            else:
                pa, pb, pc = abs(b - cc), abs(a - cc), abs(a + b - 2 * cc)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
            out[i] = (line[i] - pred) & 255
        raw += bytes([f]) + out.astype(np.uint8).tobytes()
        prev = line

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, {3: 2, 4: 6}[c], 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b""))


def _write_scene(tmp_path):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size=(12, 10, 3), dtype=np.uint8)
    (tmp_path / "test.obj").write_text(OBJ)
    (tmp_path / "test.mtl").write_text(MTL)
    (tmp_path / "floor.png").write_bytes(_png_with_filters(img))
    return img


def test_png_round_trip_all_filters():
    rng = np.random.default_rng(0)
    for c in (3, 4):
        img = rng.integers(0, 256, size=(11, 7, c), dtype=np.uint8)
        assert np.array_equal(ingest.decode_png(_png_with_filters(img)), img)
        assert np.array_equal(ingest.decode_png(ingest.encode_png(img)), img)
    lin = ingest.srgb_to_linear(np.array([0, 10, 128, 255], dtype=np.uint8))
    assert lin[0] == 0 and lin[3] == 1 and abs(lin[2] - 0.2158605) < 1e-6 and abs(lin[1] - 10 / 255 / 12.92) < 1e-7


def test_obj_mtl_conventions(tmp_path):
    img = _write_scene(tmp_path)
    scene = ingest.load_obj_scene(str(tmp_path / "test.obj"), camera_position=(0.0, 2.0, 6.0))
    assert [m.triangles.shape[0] for m in scene.meshes] == [2, 4, 2]                  # Triangulate: fans
    floor, pyramid, lamp = scene.meshes
    assert floor.positions.shape[0] == 4                                               # JoinIdenticalVertices
    assert pyramid.positions.shape[0] == 12 and lamp.positions.shape[0] == 6          # GenNormals: per-face vertices
    assert np.allclose(floor.normals, [0, 1, 0])
    n = np.cross(pyramid.positions[1] - pyramid.positions[0], pyramid.positions[2] - pyramid.positions[0])
    assert np.allclose(pyramid.normals[0], n / np.linalg.norm(n), atol=1e-6)
    assert np.allclose(sorted(floor.texcoords[:, 1].tolist()), [-1, -1, 1, 1])        # FlipUVs: v -> 1 - v
    assert np.allclose(floor.tangents, [1, 0, 0], atol=1e-6)                          # CalcTangentSpace: dP/du
    mats = scene.materials
    assert mats[0]["bsdfType"] == scenes.BSDF_DIFFUSE_AND_SPECULAR
    assert np.isclose(mats[0]["p2"], math.sqrt(25.0) / 11.0) and np.isclose(mats[1]["p2"], 10.0 / 11.0)  # sqrt(Ns) / 11
    assert mats[2]["hasEmittance"] == 1 and np.allclose(mats[2]["emittance"], [30, 28, 25])
    assert scene.material_textures[0, 0] == 0 and (scene.material_textures[1:] == 0xFFFFFFFF).all()
    tex = scene.textures[0]
    assert tex.shape == (12, 10, 4) and np.allclose(tex[..., :3], ingest.srgb_to_linear(img)) and (tex[..., 3] == 1).all()


def test_oracle_renders_the_ingested_scene_with_its_texture(tmp_path):
    _write_scene(tmp_path)
    scene = ingest.load_obj_scene(str(tmp_path / "test.obj"), camera_position=(0.0, 2.0, 6.0),
                                  camera_orientation=(scenes.rot_y(180.0) @ scenes.rot_x(15.0)))
    w, h = 48, 32
    osc = O.OracleScene(scene)
    fr = O.OracleFrame(osc, w, h)
    p = abi.default_frame_params(scene, w, h)
    for f in range(2):
        p.numAccumFrames = f
        for kind, pid in engine.restir_frame_passes(p, f, 1):
            fr.gbuffer(p) if kind == "gbuffer" else fr.restir(p, pid)
    gb0 = fr.buffer(abi.BUF_GBUFFER0, p.bufferIndex)
    floor_px = gb0[..., 1] == 0
    assert floor_px.sum() > 100
    albedo = fr.buffer(abi.BUF_ALBEDO_ACCUM)[..., :3]
    assert len(np.unique(albedo[floor_px].round(4), axis=0)) > 50, "the floor's albedo must vary with the texture"
    beauty = fr.buffer(abi.BUF_BEAUTY_ACCUM)[..., :3]
    assert np.isfinite(beauty).all() and beauty[floor_px].mean() > 1e-3
    # a textured emitter is refused
    (tmp_path / "test.mtl").write_text(MTL + "map_Ke floor.png\n")
    try:
        ingest.load_obj_scene(str(tmp_path / "test.obj"))
        raise AssertionError("map_Ke must be refused")
    except ValueError:
        pass
