"""CPU checks of the oracle's instance update (orc_scene_update_instances, the restatement of the effect of
InstanceController::update + Scene::updateASs on the hot path): it must be equivalent to building the moved scene from
scratch, a round trip must restore the original frame, and a moved object must show up in the motion vectors."""
import numpy as np

from gfxexp_b200 import abi, engine, scenes


def _render(oracle, oframe, p, frame):
    for kind, pass_id in engine.restir_frame_passes(p, frame, 1, temporal=True, unbiased=False):
        if kind == "gbuffer":
            oframe.gbuffer(p)
        else:
            oframe.restir(p, pass_id)


def _moved(scene, k=1.0):
    moved = list(scene.instances)
    for i in (1, 2):
        t = np.asarray(scene.instances[i].transform, dtype=np.float64)
        scale = float(np.linalg.norm(t[:, 0]))
        yaw = float(np.degrees(np.arctan2(t[0, 2], t[0, 0])))
        moved[i] = scenes.move_instance(scene.instances[i], translate=(t[0, 3] + 0.3 * k, t[1, 3], t[2, 3] - 0.2 * k),
                                        yaw_deg=yaw + 12.0 * k, scale=scale * (1.0 + 0.1 * k))
    return moved


def test_update_equals_fresh_build(oracle):
    scene = scenes.tiny_city_scene()
    w, h = 48, 28
    moved = _moved(scene)
    # (a) update in place
    a = oracle.OracleScene(scene)
    a.update_instances(abi.make_instance_descs(moved))
    # (b) the moved scene built from scratch
    scene_b = scenes.tiny_city_scene()
    scene_b.instances = moved
    b = oracle.OracleScene(scene_b)
    assert a.validate() == "" and b.validate() == ""
    na, ra, ta = a.export_bvh()
    nb, rb, tb = b.export_bvh()
    assert np.array_equal(ta.view(np.uint8), tb.view(np.uint8)) and np.array_equal(ra, rb)
    assert np.array_equal(na.view(np.uint8), nb.view(np.uint8))
    wa, ca, ia = a.light_dist()
    wb, cb, ib = b.light_dist()
    assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32)) and np.float32(ia).view(np.uint32) == np.float32(ib).view(np.uint32)
    # same first frame (curToPrevTransform only matters for the motion vectors, which both sides share)
    fa, fb = oracle.OracleFrame(a, w, h), oracle.OracleFrame(b, w, h)
    pa, pb = abi.default_frame_params(scene, w, h), abi.default_frame_params(scene, w, h)
    pa.log2NumCandidateSamples = pb.log2NumCandidateSamples = 2
    _render(oracle, fa, pa, 0)
    _render(oracle, fb, pb, 0)
    for buf in (abi.BUF_GBUFFER0, abi.BUF_GBUFFER1, abi.BUF_GBUFFER2, abi.BUF_BEAUTY_ACCUM):
        assert np.array_equal(fa.buffer(buf, 0).view(np.uint32), fb.buffer(buf, 0).view(np.uint32)), buf


def test_round_trip_and_motion_vectors(oracle):
    scene = scenes.tiny_city_scene()
    w, h = 48, 28
    osc = oracle.OracleScene(scene)
    p = abi.default_frame_params(scene, w, h)
    p.log2NumCandidateSamples = 2
    reference = oracle.OracleFrame(osc, w, h)
    _render(oracle, reference, p, 0)
    static_mv = reference.buffer(abi.BUF_GBUFFER1, 0).copy()
    assert not static_mv.any()                                   # static scene, static camera: no motion

    moved = _moved(scene)
    osc.update_instances(abi.make_instance_descs(moved))
    _render(oracle, reference, p, 1)                             # second frame of the same sequence (frame 0 resets the flow)
    mv = reference.buffer(abi.BUF_GBUFFER1, p.bufferIndex)
    moving_px = np.abs(mv).sum(-1) > 1e-4
    assert moving_px.any() and not moving_px.all()               # the movers move, the ground and the other buildings do not
    assert np.isfinite(mv).all()
    # a fresh sequence on the moved scene gives a different picture than the original placement
    ref0 = oracle.OracleFrame(oracle.OracleScene(scene), w, h)
    p0 = abi.default_frame_params(scene, w, h)
    p0.log2NumCandidateSamples = 2
    _render(oracle, ref0, p0, 0)
    f1 = oracle.OracleFrame(osc, w, h)
    p1 = abi.default_frame_params(scene, w, h)
    p1.log2NumCandidateSamples = 2
    _render(oracle, f1, p1, 0)
    assert not np.array_equal(f1.buffer(abi.BUF_GBUFFER0, 0), ref0.buffer(abi.BUF_GBUFFER0, 0))

    # back to the original placement (curToPrev of the way back is irrelevant for a fresh first frame's shading)
    def to44(m34):
        m = np.eye(4, dtype=np.float64)
        m[:3, :] = np.asarray(m34, dtype=np.float64)
        return m
    back = list(moved)
    for i in (1, 2):  # the original placement, bit for bit, with the matching curToPrevTransform
        orig = scene.instances[i]
        cur_to_prev = (to44(moved[i].transform) @ np.linalg.inv(to44(orig.transform)))[:3, :].astype(np.float32)
        back[i] = scenes.Instance(orig.transform, cur_to_prev, orig.normal_matrix, orig.uniform_scale, list(orig.mesh_slots))
    osc.update_instances(abi.make_instance_descs(back))
    f2 = oracle.OracleFrame(osc, w, h)
    p2 = abi.default_frame_params(scene, w, h)
    p2.log2NumCandidateSamples = 2
    _render(oracle, f2, p2, 0)
    for buf in (abi.BUF_GBUFFER0, abi.BUF_GBUFFER2, abi.BUF_GBUFFER3, abi.BUF_BEAUTY_ACCUM, abi.BUF_RESERVOIR):
        assert np.array_equal(f2.buffer(buf, 0).view(np.uint32), ref0.buffer(buf, 0).view(np.uint32)), buf
