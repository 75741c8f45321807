"""CPU checks of the oracle's environment light that do not need the reference binary:
  - the importance map is a normalised density, sample() and evaluatePDF() agree, and samples follow luminance x sin(theta);
  - the software tex2DLod returns the texel at texel centres and the exact mean half-way between two centres;
  - white-furnace pin: a Lambert plane of albedo rho under a uniform environment of radiance L radiates rho x L.  The
    ReSTIR DI estimator (importance-map candidates, the 2 pi^2 sin(theta) Jacobian of restir_di_shared.h:348-355) and the path
    tracer (NEE + the miss program's implicit hit, MIS-weighted) must both average to it."""
import math

import numpy as np

from gfxexp_b200 import abi, engine, scenes
from tests import oracle_lib as O


def _sky_scene(size=(64, 32)):
    scene = scenes.tiny_city_scene()
    scene.env_map = scenes.procedural_sky(*size)
    return scene


def test_importance_map_is_a_density_and_sampling_follows_it():
    scene = _sky_scene()
    osc = O.OracleScene(scene, build_bvh=False)
    H, W = scene.env_map.shape[:2]
    centres = np.stack(np.meshgrid((np.arange(W) + 0.5) / W, (np.arange(H) + 0.5) / H), -1).reshape(-1, 2).astype(np.float32)
    pdf = osc.env_query(1, centres)[:, 0].reshape(H, W)
    assert abs(pdf.mean() - 1.0) < 1e-5                       # integrates to one over the unit square
    lum = (scene.env_map[..., :3] * np.array([0.2126729, 0.7151522, 0.0721750], dtype=np.float32)).sum(-1)
    want = lum * np.sin(np.pi * (np.arange(H) + 0.5) / H)[:, None]
    assert np.allclose(pdf / pdf.sum(), want / want.sum(), rtol=2e-5)
    rng = np.random.default_rng(1)
    u = rng.random((400000, 2), dtype=np.float32)
    smp = osc.env_query(0, u)
    assert smp[:, :2].min() >= 0.0 and smp[:, :2].max() < 1.0
    # the returned density is the density at the sample (a sample that rounds onto a cell border may read the next cell)
    assert (osc.env_query(1, smp[:, :2])[:, 0] != smp[:, 2]).mean() < 1e-4
    ix = np.minimum((smp[:, 0] * W).astype(int), W - 1)
    iy = np.minimum((smp[:, 1] * H).astype(int), H - 1)
    hist = np.zeros((H, W))
    np.add.at(hist, (iy, ix), 1)
    expect = pdf / pdf.sum() * len(u)
    big = expect > 50
    chi2 = ((hist[big] - expect[big]) ** 2 / expect[big]).sum() / big.sum()  # ~1 for samples that follow the density
    assert 0.8 < chi2 < 1.2, chi2
    assert np.abs(hist[big] - expect[big]).max() < 6 * np.sqrt(expect[big].max())


def test_software_texture_fetch():
    scene = _sky_scene((8, 4))
    osc = O.OracleScene(scene, build_bvh=False)
    H, W = 4, 8
    tex = scene.env_map[..., :3]
    centres = np.stack(np.meshgrid((np.arange(W) + 0.5) / W, (np.arange(H) + 0.5) / H), -1).reshape(-1, 2).astype(np.float32)
    assert np.array_equal(osc.env_query(2, centres).reshape(H, W, 3), tex)
    mid = np.array([[2.0 / W, 0.5 / H]], dtype=np.float32)           # half-way between texels (1, 0) and (2, 0)
    assert np.allclose(osc.env_query(2, mid)[0], 0.5 * (tex[0, 1] + tex[0, 2]), rtol=1e-6)
    corner = np.array([[-3.0, 7.0]], dtype=np.float32)                # clamp addressing
    assert np.array_equal(osc.env_query(2, corner)[0], tex[H - 1, 0])


def _furnace_scene(radiance, albedo):
    mats = np.zeros(1, dtype=scenes.MATERIAL_DTYPE)
    mats[0]["p0"] = np.float32(albedo)
    mats[0]["bsdfType"] = scenes.BSDF_LAMBERT
    plane = scenes.make_quad_light(4000.0, 0)                         # a quad in the xz-plane facing -y, not emissive here
    ori = scenes.rot_x(-90.0).astype(np.float32)                      # the camera looks along +y, up at the quad
    scene = scenes.Scene([plane], mats, [scenes.make_instance([0])], np.array([0.0, -100.0, 0.0], dtype=np.float32), ori,
                         math.radians(50.0), "furnace")
    env = np.ones((16, 32, 4), dtype=np.float32)
    env[..., :3] = radiance
    scene.env_map = env
    return scene


def test_white_furnace_restir_and_path_tracer():
    L, rho = 2.0, 0.5
    scene = _furnace_scene(L, rho)
    w = h = 24
    osc = O.OracleScene(scene)
    p = abi.default_frame_params(scene, w, h)
    p.envLightPowerCoeff = 1.5
    want = rho * L * 1.5
    # ReSTIR DI, initial candidates only, accumulated over frames
    fr = O.OracleFrame(osc, w, h)
    p.log2NumCandidateSamples = 3
    for f in range(12):
        p.numAccumFrames = f
        for kind, pid in engine.restir_frame_passes(p, f, 0, temporal=False):
            fr.gbuffer(p) if kind == "gbuffer" else fr.restir(p, pid)
    assert (fr.buffer(abi.BUF_GBUFFER0, p.bufferIndex)[..., 0] != 0xFFFFFFFF).all(), "the camera must see the plane only"
    beauty = fr.buffer(abi.BUF_BEAUTY_ACCUM)[..., :3]
    assert abs(beauty.mean() / want - 1) < 0.02, beauty.mean()
    # path tracer: NEE + implicit environment hits through the miss program, MIS
    fr2 = O.OracleFrame(osc, w, h)
    p.maxPathLength = 5
    for f in range(12):
        p.numAccumFrames = f
        p.frameIndex, p.bufferIndex = f, f % 2
        fr2.gbuffer(p)
        fr2.pathtrace(p)
    beauty = fr2.buffer(abi.BUF_BEAUTY_ACCUM)[..., :3]
    assert abs(beauty.mean() / want - 1) < 0.02, beauty.mean()
    # a pixel that sees the sky shows envLightPowerCoeff x the texel
    scene.camera_orientation = scenes.rot_x(90.0).astype(np.float32)  # look down, away from the plane
    osc2 = O.OracleScene(scene)
    p2 = abi.default_frame_params(scene, w, h)
    p2.envLightPowerCoeff = 1.5
    fr3 = O.OracleFrame(osc2, w, h)
    for kind, pid in engine.restir_frame_passes(p2, 0, 0, temporal=False):
        fr3.gbuffer(p2) if kind == "gbuffer" else fr3.restir(p2, pid)
    assert np.allclose(fr3.buffer(abi.BUF_BEAUTY_ACCUM)[..., :3], 1.5 * L, rtol=1e-6)
