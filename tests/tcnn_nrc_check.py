#!/usr/bin/env python
"""Pins the NRC network against the reference's own code: tiny-cuda-nn's NetworkWithInputEncoding +
Trainer built from neural_radiance_caching/network_interface.cu's config, compiled from /root/reference/ext/tiny-cuda-nn where it
lies into oracle/_ref/libtcnn_nrc.so (oracle/ref_tcnn/tcnn_nrc.cu, Makefile).  Runs on a GPU box in its own process (third-party
kernels: a fault there must not take the test session's CUDA context with it) and prints ONE JSON line:

  init        tcnn's pcg32{1337} initial parameters vs gfx_nrc_create's (bit-exact fp32 master weights wanted)
  forward     relative L2 of the inferred radiance: this repo's tcgen05 kernel / the oracle (fp32 and half accumulation) vs tcnn
  gradients   one training step from identical weights: relative L2 per weight matrix and for the hash grid, same three arms
  weights     training / inference (EMA) weights after four training steps
  timing      ms per call of tcnn's own launches and of this repo's, same B200, same inputs (2 116 608 queries; 16 384 samples)

Usage: python tests/tcnn_nrc_check.py [--no-timing]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle_lib as O
from gfxexp_b200 import abi, engine

HIDDEN, LR = 2, 1e-2  # neural_radiance_caching_main.cpp:458-460


class TcnnNrc:
    def __init__(self):
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtcnn_nrc.so"))
        vp, u32 = C.c_void_p, C.c_uint32
        L = self.lib
        L.tcnn_nrc_last_error.restype = C.c_char_p
        L.tcnn_nrc_create.argtypes = [u32, u32, C.c_float, C.POINTER(vp)]
        L.tcnn_nrc_destroy.argtypes = [vp]
        L.tcnn_nrc_num_params.argtypes = [vp]
        L.tcnn_nrc_num_params.restype = u32
        L.tcnn_nrc_read.argtypes = [vp, C.c_int, vp]
        L.tcnn_nrc_set_params.argtypes = [vp, vp]
        L.tcnn_nrc_infer.argtypes = [vp, vp, u32, vp]
        L.tcnn_nrc_train.argtypes = [vp, vp, vp, u32, C.POINTER(C.c_float)]
        L.tcnn_nrc_forward_backward.argtypes = [vp, vp, vp, u32, C.POINTER(C.c_float), vp]
        L.tcnn_nrc_time_infer.argtypes = [vp, vp, u32, u32, u32, C.POINTER(C.c_float)]
        L.tcnn_nrc_time_train.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(C.c_float)]
        self.h = vp()
        self._ok(L.tcnn_nrc_create(1, HIDDEN, LR, C.byref(self.h)))  # 1 = PositionEncoding::HashGrid
        self.num_params = L.tcnn_nrc_num_params(self.h)

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError("tiny-cuda-nn: " + self.lib.tcnn_nrc_last_error().decode())

    def read(self, which):
        out = np.empty(self.num_params, dtype=np.float32 if which == 0 else np.float16)
        self._ok(self.lib.tcnn_nrc_read(self.h, which, out.ctypes.data))
        return out

    def set_params(self, master):
        m = np.ascontiguousarray(master, dtype=np.float32)
        self._ok(self.lib.tcnn_nrc_set_params(self.h, m.ctypes.data))

    def infer(self, q):
        out = np.empty((q.shape[0], 3), dtype=np.float32)
        self._ok(self.lib.tcnn_nrc_infer(self.h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    def train(self, q, t):
        loss = C.c_float()
        self._ok(self.lib.tcnn_nrc_train(self.h, q.ctypes.data, t.ctypes.data, q.shape[0], C.byref(loss)))
        return loss.value

    def forward_backward(self, q, t):
        loss = C.c_float()
        self._ok(self.lib.tcnn_nrc_forward_backward(self.h, q.ctypes.data, t.ctypes.data, q.shape[0], C.byref(loss), None))
        return loss.value, self.read(3).astype(np.float32)

    def time_infer(self, q, warmup, iters):
        ms = C.c_float()
        self._ok(self.lib.tcnn_nrc_time_infer(self.h, q.ctypes.data, q.shape[0], warmup, iters, C.byref(ms)))
        return ms.value

    def time_train(self, q, t, warmup, iters):
        ms = C.c_float()
        self._ok(self.lib.tcnn_nrc_time_train(self.h, q.ctypes.data, t.ctypes.data, q.shape[0], warmup, iters, C.byref(ms)))
        return ms.value


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def queries(rng, n):
    """radiance queries as createRadianceQuery writes them (optix_pathtracing_kernels.cu:12-34): position in the unit cube,
    polar angles, 1 - exp(-roughness), reflectances - everything in [0, 1] but the angles, which span [-pi, pi] / [0, pi]"""
    q = rng.uniform(0.0, 1.0, size=(n, 14)).astype(np.float32)
    q[:, 3] = rng.uniform(-np.pi, np.pi, n)
    q[:, 5] = rng.uniform(-np.pi, np.pi, n)
    q[:, 4] = rng.uniform(0.0, np.pi, n)
    q[:, 6] = rng.uniform(0.0, np.pi, n)
    return np.ascontiguousarray(q)


def targets(q):
    """a smooth radiance field of the query, so that training has something to learn"""
    t = np.stack([0.5 + 0.5 * np.sin(6 * q[:, 0] + q[:, 8]), q[:, 1] * q[:, 9] + 0.1, np.abs(np.cos(4 * q[:, 2])) * q[:, 10]], 1)
    return np.ascontiguousarray(t.astype(np.float32) * 2.0)


def layer_slices(num_matrix_weights, num_params):
    s, out = 0, {}
    for l in range(HIDDEN):
        out[f"W{l}"] = slice(s, s + 64 * 64)
        s += 64 * 64
    out[f"W{HIDDEN}"] = slice(s, s + 16 * 64)
    out["grid"] = slice(num_matrix_weights, num_params)
    return out


def main():
    timing = "--no-timing" not in sys.argv
    result = {}
    ref = TcnnNrc()
    ctx = engine.Context(0)
    ours = engine.NeuralRadianceCache(ctx, HIDDEN, LR)
    assert ours.num_params == ref.num_params, (ours.num_params, ref.num_params)
    P = ours.num_params
    onet = O.OracleNrc(HIDDEN, LR)
    M = onet.num_matrix_weights
    slices = layer_slices(M, P)

    # ---- init: Trainer{seed 1337}
    ref_master, our_master = ref.read(0), ours.read(abi.NRC_READ_MASTER)
    mism = ref_master.view(np.uint32) != our_master.view(np.uint32)
    result["init"] = {"params": int(P), "master_mismatches": int(mism.sum()), "mlp_mismatches": int(mism[:M].sum()),
                      "max_abs_diff": float(np.abs(ref_master - our_master).max()),
                      "training_half_mismatches": int((ref.read(1).view(np.uint16) != ours.read(abi.NRC_READ_TRAINING).view(np.uint16)).sum()),
                      "inference_all_zero": bool(not ref.read(2).any() and not ours.read(abi.NRC_READ_INFERENCE).any())}

    # ---- forward and one training step from identical, half-representable weights of working magnitude
    rng = np.random.default_rng(23)
    params = engine.random_nrc_params(P, M, seed=31, grid_amplitude=0.5)      # float16
    master = params.astype(np.float32)
    ref.set_params(master)
    ours.set_params(params)
    onet.set_params(params)
    n_inf = 128 * 256
    q = queries(rng, n_inf)
    ref_out = ref.infer(q)
    dq = torch.from_numpy(q).cuda()
    dout = torch.empty((n_inf, 3), device="cuda")
    ours.infer(dq, dout, n_inf)
    torch.cuda.synchronize()
    our_out = dout.cpu().numpy()
    o32 = onet.infer(q)
    onet.set_accumulate_half(True)
    o16 = onet.infer(q)
    onet.set_accumulate_half(False)
    result["forward"] = {"n": n_inf, "ours_vs_tcnn": rel_l2(our_out, ref_out), "oracle_fp32acc_vs_tcnn": rel_l2(o32, ref_out),
                         "oracle_halfacc_vs_tcnn": rel_l2(o16, ref_out), "ours_vs_oracle_fp32acc": rel_l2(our_out, o32),
                         "tcnn_rms": float(np.sqrt(np.mean(ref_out.astype(np.float64) ** 2)))}

    n_tr = 16384
    qt = queries(rng, n_tr)
    tt = targets(qt)
    ref_loss, ref_grad = ref.forward_backward(qt, tt)
    ours.keep_gradients(True)
    dqt, dtt = torch.from_numpy(qt).cuda(), torch.from_numpy(tt).cuda()
    our_loss = ours.train(dqt, dtt, n_tr, want_loss=True)
    our_grad = ours.read(abi.NRC_READ_GRADIENTS)
    o_loss32 = onet.train(qt, tt)
    o_grad32 = onet.get_gradients()
    onet.set_params(params)
    onet.set_accumulate_half(True)
    o_loss16 = onet.train(qt, tt)
    o_grad16 = onet.get_gradients()
    onet.set_accumulate_half(False)
    grads = {"loss": {"tcnn": ref_loss, "ours": our_loss, "oracle_fp32acc": o_loss32, "oracle_halfacc": o_loss16}}
    for name, sl in slices.items():
        grads[name] = {"ours_vs_tcnn": rel_l2(our_grad[sl], ref_grad[sl]), "oracle_fp32acc_vs_tcnn": rel_l2(o_grad32[sl], ref_grad[sl]),
                       "oracle_halfacc_vs_tcnn": rel_l2(o_grad16[sl], ref_grad[sl]), "ours_vs_oracle_fp32acc": rel_l2(our_grad[sl], o_grad32[sl]),
                       "tcnn_norm": float(np.linalg.norm(ref_grad[sl].astype(np.float64))),
                       "tcnn_nonzero": int(np.count_nonzero(ref_grad[sl]))}
    result["gradients"] = grads

    # ---- four full training steps (forward, loss, backward, Adam, EMA) from the same start
    ref2 = TcnnNrc()
    ref2.set_params(master)
    ours.set_params(params)
    onet.set_params(params)
    losses = {"tcnn": [], "ours": [], "oracle_fp32acc": []}
    for step in range(4):
        qs = queries(rng, n_tr)
        ts = targets(qs)
        losses["tcnn"].append(ref2.train(qs, ts))
        losses["ours"].append(ours.train(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), n_tr, want_loss=True))
        losses["oracle_fp32acc"].append(onet.train(qs, ts))
    weights = {"losses": losses}
    ref_tr, ref_inf = ref2.read(1).astype(np.float32), ref2.read(2).astype(np.float32)
    our_tr, our_inf = ours.read(abi.NRC_READ_TRAINING).astype(np.float32), ours.read(abi.NRC_READ_INFERENCE).astype(np.float32)
    orc_tr, orc_inf = onet.get_params(ema=False).astype(np.float32), onet.get_params(ema=True).astype(np.float32)
    delta_ref = ref_tr - master       # what four steps changed: the comparison that matters is on the update, not on the weights
    for name, sl in slices.items():
        weights[name] = {"training_ours_vs_tcnn": rel_l2(our_tr[sl], ref_tr[sl]), "inference_ours_vs_tcnn": rel_l2(our_inf[sl], ref_inf[sl]),
                         "training_oracle_vs_tcnn": rel_l2(orc_tr[sl], ref_tr[sl]), "inference_oracle_vs_tcnn": rel_l2(orc_inf[sl], ref_inf[sl]),
                         "update_ours_vs_tcnn": rel_l2(our_tr[sl] - master[sl], delta_ref[sl]),
                         "update_oracle_vs_tcnn": rel_l2(orc_tr[sl] - master[sl], delta_ref[sl])}
    qe = queries(rng, n_inf)
    ref_after = ref2.infer(qe)
    dqe = torch.from_numpy(qe).cuda()
    ours.infer(dqe, dout, n_inf)
    torch.cuda.synchronize()
    weights["inference_after_training_ours_vs_tcnn"] = rel_l2(dout.cpu().numpy(), ref_after)
    weights["inference_after_training_oracle_vs_tcnn"] = rel_l2(onet.infer(qe), ref_after)
    result["weights"] = weights

    # ---- timing: the reference's own kernels and this repo's on the same GPU, same inputs
    if timing:
        n_big = 2_116_608  # pad128(1920 * 1080 + 1920 * 1080 / 48): a 1080p NRC frame's query count
        qb = queries(rng, n_big)
        t_ref_inf = ref2.time_infer(qb, 3, 20)
        dqb = torch.from_numpy(qb).cuda()
        doutb = torch.empty((n_big, 3), device="cuda")
        for _ in range(3):
            ours.infer(dqb, doutb, n_big)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ours.infer(dqb, doutb, n_big)
        e1.record()
        torch.cuda.synchronize()
        t_our_inf = e0.elapsed_time(e1) / 20
        t_ref_tr = ref2.time_train(qt, tt, 3, 40)
        for _ in range(3):
            ours.train(dqt, dtt, n_tr)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(40):
            ours.train(dqt, dtt, n_tr)
        e1.record()
        torch.cuda.synchronize()
        t_our_tr = e0.elapsed_time(e1) / 40
        flop_inf, flop_tr = 18432.0 * n_big, 55296.0 * n_tr
        result["timing"] = {"infer_queries": n_big, "tcnn_infer_ms": t_ref_inf, "ours_infer_ms": t_our_inf,
                            "tcnn_infer_tflops": flop_inf / t_ref_inf / 1e9, "ours_infer_tflops": flop_inf / t_our_inf / 1e9,
                            "train_samples": n_tr, "tcnn_train_step_ms": t_ref_tr, "ours_train_step_ms": t_our_tr,
                            "tcnn_train_tflops": flop_tr / t_ref_tr / 1e9, "ours_train_tflops": flop_tr / t_our_tr / 1e9}
    print(json.dumps(result))
    return 0


if __name__ == "__main__":
    sys.exit(main())
