"""GPU parity: LBVH build + traversal vs the CPU oracle."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, scenes

pytestmark = pytest.mark.gpu


def _hit_mismatch(a, b):
    return np.nonzero((a["dist"].view(np.uint32) != b["dist"].view(np.uint32)) | (a["primIndex"] != b["primIndex"])
                      | (a["geomIndex"] != b["geomIndex"]) | (a["instIndex"] != b["instIndex"])
                      | ((a["bcB"].view(np.uint32) != b["bcB"].view(np.uint32)) & (a["primIndex"] != 0xFFFFFFFF))
                      | ((a["bcC"].view(np.uint32) != b["bcC"].view(np.uint32)) & (a["primIndex"] != 0xFFFFFFFF)))[0]


def _random_rays(info, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(info.sceneMin), np.array(info.sceneMax)
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["org"] = (lo + (hi - lo) * rng.uniform(-0.2, 1.2, size=(n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"] = 0.0
    rays["tmax"] = np.float32(3.402823466e+38)
    return rays


BVH_BUILD_FAST = 0x100  # GFX_BVH_BUILD_FAST: Karras LBVH hierarchy
BVH_BUILD_PLOC = 0x200  # GFX_BVH_BUILD_PLOC: PLOC clustering; the default (0) is the top-down binned-SAH hierarchy


@pytest.mark.parametrize("scene_fn,w,h,flags", [(scenes.tiny_city_scene, 160, 96, 0), (scenes.small_city_scene, 320, 200, 0),
                                                (scenes.small_city_scene, 320, 200, BVH_BUILD_FAST),
                                                (scenes.small_city_scene, 320, 200, BVH_BUILD_PLOC),
                                                (scenes.tiny_city_scene, 160, 96, BVH_BUILD_PLOC | 4),
                                                (scenes.tiny_city_scene, 160, 96, 1), (scenes.small_city_scene, 320, 200, 8),
                                                (scenes.tiny_city_scene, 160, 96, BVH_BUILD_FAST | 2)])
def test_lbvh_format_and_closest_hit(gfx_ctx, oracle, scene_fn, w, h, flags):
    scene = scene_fn()
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh(flags)
    info = gfx_ctx.bvh_info()
    assert info.numTriangles == scene.num_triangles
    assert info.numPrimRefs == info.numTriangles  # LBVH never duplicates references

    oscene = oracle.OracleScene(scene)
    # (1) the GPU-built BVH is a valid reference-format BVH: the restated bvh::traverse walks it
    nodes, refs, tris = gfx_ctx.export_bvh()
    onodes, orefs, otris = oscene.export_bvh()
    assert np.array_equal(tris.view(np.uint8), otris.view(np.uint8)), "TriangleStorage differs from the oracle's"
    gpu_bvh_on_cpu = oracle.OracleScene(scene)
    gpu_bvh_on_cpu.import_bvh(nodes, refs, tris)
    assert gpu_bvh_on_cpu.validate() == ""

    p = abi.default_frame_params(scene, w, h)
    rays = np.concatenate([oracle.primary_rays(p, w, h), _random_rays(info, 20000, 7)])
    want = oscene.trace(rays, oracle.TRACE_CANONICAL)
    via_gpu_bvh_cpu_traverse = gpu_bvh_on_cpu.trace(rays, oracle.TRACE_CANONICAL)
    assert len(_hit_mismatch(via_gpu_bvh_cpu_traverse, want)) == 0

    # (2) GPU traversal of the GPU LBVH == oracle (bit-exact dist / ids / barycentrics)
    got = gfx_ctx.trace(rays, abi.TRACE_CLOSEST)
    bad = _hit_mismatch(got, want)
    assert len(bad) == 0, f"{len(bad)} of {len(rays)} rays differ, first: {got[bad[:3]]} vs {want[bad[:3]]}"

    # (3) GPU traversal of the oracle-built SBVH (reference builder) gives the same hits
    gfx_ctx.import_bvh(onodes, orefs, otris)
    got2 = gfx_ctx.trace(rays, abi.TRACE_CLOSEST)
    assert len(_hit_mismatch(got2, want)) == 0


def test_brute_force_agreement(gfx_ctx, oracle):
    scene = scenes.tiny_city_scene()
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    oscene = oracle.OracleScene(scene)
    rays = _random_rays(gfx_ctx.bvh_info(), 4096, 11)
    want = oscene.trace(rays, oracle.TRACE_BRUTE_FORCE)
    got = gfx_ctx.trace(rays, abi.TRACE_CLOSEST)
    assert len(_hit_mismatch(got, want)) == 0


def test_any_hit(gfx_ctx, oracle):
    scene = scenes.small_city_scene()
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh()
    oscene = oracle.OracleScene(scene)
    rays = _random_rays(gfx_ctx.bvh_info(), 50000, 3)
    rays["tmax"] = np.random.default_rng(5).uniform(0.5, 30.0, size=len(rays)).astype(np.float32)
    want = oscene.trace(rays, oracle.TRACE_ANY)
    got = gfx_ctx.trace(rays, abi.TRACE_ANY)
    assert np.array_equal(got["primIndex"] == 0xFFFFFFFF, want["primIndex"] == 0xFFFFFFFF)
    assert np.array_equal(got["dist"], want["dist"])


def test_edge_cases(gfx_ctx, oracle):
    # single triangle, degenerate (zero-area) triangle, rays parallel to an axis, empty ray batch
    mats = np.zeros(1, dtype=scenes.MATERIAL_DTYPE)
    mats[0]["bsdfType"] = scenes.BSDF_LAMBERT
    mats[0]["p0"] = 0.5
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 2, 2]], dtype=np.float32)
    for faces in (np.array([[0, 1, 2]], dtype=np.uint32), np.array([[0, 1, 2], [3, 3, 3]], dtype=np.uint32)):
        mesh = scenes.mesh_from_triangles(verts, faces, 0)
        sc = scenes.Scene([mesh], mats, [scenes.make_instance([0])], np.zeros(3, np.float32), np.eye(3, dtype=np.float32), 1.0)
        gfx_ctx.upload_scene(sc)
        gfx_ctx.build_bvh()
        oscene = oracle.OracleScene(sc)
        rays = np.zeros(4, dtype=abi.RAY_DTYPE)
        rays["org"] = [[0.25, 0.25, 1], [0.25, 0.25, 1], [5, 5, 1], [0.25, 0.25, -1]]
        rays["dir"] = [[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, 1]]
        rays["tmax"] = 100.0
        got = gfx_ctx.trace(rays)
        want = oscene.trace(rays, oracle.TRACE_BRUTE_FORCE)
        assert len(_hit_mismatch(got, want)) == 0
        assert got["primIndex"][0] == 0 and got["primIndex"][1] == 0xFFFFFFFF
        assert len(gfx_ctx.trace(np.zeros(0, dtype=abi.RAY_DTYPE))) == 0


def test_launch_before_ready_fails_loudly():
    ctx = engine.Context(0)
    with pytest.raises(engine.GfxError):
        ctx.build_bvh()
    with pytest.raises(engine.GfxError):
        ctx.trace(np.zeros(1, dtype=abi.RAY_DTYPE))
    ctx.close()
