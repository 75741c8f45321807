"""Edge cases through the C ABI, GPU vs oracle: a scene without emitters (every light pdf is 0), a single triangle, image
sizes that are not multiples of the 8x8 / 32x8 thread tiles, a camera inside geometry looking at nothing."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, scenes

pytestmark = pytest.mark.gpu


def _same(got, want, tag):
    g, w = np.ascontiguousarray(got), np.ascontiguousarray(want)
    g = g.view(np.uint32) if g.dtype.itemsize == 4 else g
    w = w.view(np.uint32) if w.dtype.itemsize == 4 else w
    assert np.array_equal(g, w), f"{tag} differs"


def _frames(ctx, oracle, scene, w, h, frames=2):
    ctx.upload_scene(scene)
    ctx.build_bvh()
    ctx.create_frame(w, h)
    oframe = oracle.OracleFrame(oracle.OracleScene(scene), w, h)
    p = abi.default_frame_params(scene, w, h)
    for f in range(frames):
        p.numAccumFrames = f
        ctx.build_light_distributions(f % 2)
        for kind, pid in engine.restir_frame_passes(p, f, 1, True, False):
            if kind == "gbuffer":
                ctx.gbuffer(p)
                oframe.gbuffer(p)
            else:
                ctx.restir(p, pid)
                oframe.restir(p, pid)
    for buf in (abi.BUF_GBUFFER0, abi.BUF_GBUFFER2, abi.BUF_RNG, abi.BUF_BEAUTY_ACCUM):
        _same(ctx.download(buf, 0), oframe.buffer(buf, 0), f"restir buffer {buf}")
    # then two path-traced samples on top
    for f in range(2):
        p.numAccumFrames = f
        ctx.gbuffer(p)
        ctx.pathtrace(p)
        oframe.gbuffer(p)
        oframe.pathtrace(p)
    _same(ctx.download(abi.BUF_BEAUTY_ACCUM), oframe.buffer(abi.BUF_BEAUTY_ACCUM), "path-traced beauty")
    _same(ctx.download(abi.BUF_RNG), oframe.buffer(abi.BUF_RNG), "rng")
    return ctx.download(abi.BUF_BEAUTY_ACCUM)[..., :3]


def test_scene_without_emitters(gfx_ctx, oracle):
    scene = scenes.tiny_city_scene()
    scene.materials = scene.materials.copy()
    scene.materials["hasEmittance"] = 0
    scene.materials["emittance"] = 0
    scene._aabb_cache = None
    assert scene.num_emissive_triangles == 0
    beauty = _frames(gfx_ctx, oracle, scene, 72, 44)
    assert np.isfinite(beauty).all()
    assert beauty.max() <= 0.001 + 1e-9  # nothing emits: only the miss background is non-zero


def test_single_triangle_and_ragged_image(gfx_ctx, oracle):
    verts = np.array([[-1, 0, 4], [1, 0, 4], [0, 1.5, 4]], dtype=np.float32)
    faces = np.array([[0, 1, 2]], dtype=np.uint32)
    mats = np.zeros(2, dtype=scenes.MATERIAL_DTYPE)
    mats[0]["p0"] = (0.7, 0.6, 0.5)
    mats[0]["bsdfType"] = scenes.BSDF_LAMBERT
    mats[1]["p0"] = (0.8, 0.8, 0.8)
    mats[1]["p1"] = (0.04, 0.04, 0.04)
    mats[1]["p2"] = 0.3
    mats[1]["bsdfType"] = scenes.BSDF_DIFFUSE_AND_SPECULAR
    mats[1]["emittance"] = (5, 5, 5)
    mats[1]["hasEmittance"] = 1
    tri = scenes.mesh_from_triangles(verts, faces, 0)
    light = scenes.make_quad_light(1.0, 1)
    scene = scenes.Scene(meshes=[tri, light], materials=mats,
                         instances=[scenes.make_instance([0]), scenes.make_instance([1], translate=(0, 3, 3))],
                         camera_position=np.array([0, 0.5, 0], dtype=np.float32),
                         camera_orientation=np.eye(3, dtype=np.float32), fov_y=np.float32(np.deg2rad(50)), name="one-triangle")
    beauty = _frames(gfx_ctx, oracle, scene, 37, 23)  # neither a multiple of 8 nor of 32
    assert np.isfinite(beauty).all() and beauty.max() > 0.001


def test_empty_view(gfx_ctx, oracle):
    """every primary ray misses: all pipelines must write their background constants and leave the RNG untouched"""
    scene = scenes.tiny_city_scene()
    scene.camera_position = np.array([0, 5000, 0], dtype=np.float32)
    scene.camera_orientation = scenes.look_at_orientation(scene.camera_position, scene.camera_position + np.array([0, 1, 0.001]))
    scene._aabb_cache = None
    beauty = _frames(gfx_ctx, oracle, scene, 40, 24)
    assert np.all(gfx_ctx.download(abi.BUF_GBUFFER0, 0)[..., 0] == 0xFFFFFFFF)
    assert np.all(beauty == np.float32(0.001))
    rng_after = gfx_ctx.download(abi.BUF_RNG)
    gfx_ctx.create_frame(40, 24)  # re-seeds
    assert np.array_equal(rng_after, gfx_ctx.download(abi.BUF_RNG))
