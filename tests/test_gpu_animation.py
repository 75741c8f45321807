"""GPU parity with animated instances (InstanceController::update, common/common_host.h:798-856, the step
before the hot path every frame): buildings and a street lamp move, rotate and scale; each frame both sides receive the new
transforms (+ curToPrevTransform, normal matrices), rebuild the acceleration structure and the light distributions, and
the ReSTIR DI frame - motion vectors of moving objects, temporal reuse across them, light triangles that move - must stay
bit-identical to the CPU oracle."""
import numpy as np
import pytest

from gfxexp_b200 import abi, engine, scenes

pytestmark = pytest.mark.gpu
BVH_BUILD_FAST = 0x100


ALL_BUFFERS = [(abi.BUF_GBUFFER0, 2), (abi.BUF_GBUFFER1, 2), (abi.BUF_GBUFFER2, 2), (abi.BUF_GBUFFER3, 2),
               (abi.BUF_RNG, 1), (abi.BUF_RESERVOIR, 2), (abi.BUF_RESERVOIR_INFO, 2), (abi.BUF_BEAUTY_ACCUM, 1),
               (abi.BUF_ALBEDO_ACCUM, 1), (abi.BUF_NORMAL_ACCUM, 1)]


def _compare_all(ctx, oframe, tag):
    for buf, count in ALL_BUFFERS:
        for idx in range(count):
            got = ctx.download(buf, idx)
            want = oframe.buffer(buf, idx)
            g = got.view(np.uint32) if got.dtype != np.uint64 else got
            w = want.view(np.uint32) if want.dtype != np.uint64 else want
            if not np.array_equal(g, w):
                bad = np.argwhere(g != w)
                raise AssertionError(f"{tag}: buffer {buf}[{idx}] differs at {len(bad)} elements, first {bad[:4].tolist()}")


def _emissive_instances(scene):
    emissive_mesh = [bool(scene.materials[m.material]["hasEmittance"]) for m in scene.meshes]
    return [i for i, inst in enumerate(scene.instances) if any(emissive_mesh[s] for s in inst.mesh_slots)]


def _placement(inst):
    """(translate, yaw, scale) of an instance made by scenes.make_instance with pitch 0"""
    t = np.asarray(inst.transform, dtype=np.float64)
    scale = float(np.linalg.norm(t[:, 0]))
    yaw = float(np.degrees(np.arctan2(t[0, 2], t[0, 0])))
    return t[:, 3].copy(), yaw, scale


@pytest.mark.parametrize("bvh_flags", [BVH_BUILD_FAST, 0])
def test_animated_instances_three_frames_bit_exact(gfx_ctx, oracle, bvh_flags):
    scene = scenes.tiny_city_scene()
    w, h = 128, 72
    gfx_ctx.upload_scene(scene)
    gfx_ctx.build_bvh(bvh_flags)
    gfx_ctx.create_frame(w, h)
    oscene = oracle.OracleScene(scene)
    oframe = oracle.OracleFrame(oscene, w, h)
    p = abi.default_frame_params(scene, w, h)
    p.log2NumCandidateSamples = 3

    lamps = _emissive_instances(scene)
    assert lamps, "the scene has no emissive instance"
    movers = [1, 2, lamps[0]]                      # two buildings and one lamp (instance 0 is the ground)
    current = list(scene.instances)
    base = {i: _placement(scene.instances[i]) for i in movers}
    for frame in range(3):
        if frame > 0:
            nxt = list(current)
            for k, i in enumerate(movers):
                t0, yaw0, s0 = base[i]
                nxt[i] = scenes.move_instance(current[i], translate=(t0[0] + 0.15 * frame * (k + 1), t0[1], t0[2] - 0.1 * frame),
                                              yaw_deg=yaw0 + 7.0 * frame * (k + 1), scale=s0 * (1.0 + 0.05 * frame))
            # instances that stopped moving would get curToPrev = identity again; here every mover moves every frame
            current = nxt
            descs = abi.make_instance_descs(current)
            gfx_ctx.update_instances(descs)
            gfx_ctx.build_bvh(bvh_flags)
            oscene.update_instances(descs)
        gfx_ctx.build_light_distributions(frame % 2)
        for kind, pass_id in engine.restir_frame_passes(p, frame, 1, temporal=True, unbiased=False):
            if kind == "gbuffer":
                gfx_ctx.gbuffer(p)
                oframe.gbuffer(p)
            else:
                gfx_ctx.restir(p, pass_id)
                oframe.restir(p, pass_id)
        gfx_ctx.synchronize()
        _compare_all(gfx_ctx, oframe, f"animated frame {frame} flags={bvh_flags:#x}")
    # the movers are visible: the motion-vector plane is not identically the static-scene one
    mv = gfx_ctx.download(abi.BUF_GBUFFER1, p.bufferIndex)
    assert np.isfinite(mv).all()
