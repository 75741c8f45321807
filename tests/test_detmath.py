"""detmath.h (shared deterministic transcendentals) bounded against libm on the CPU."""
import ctypes as C

import numpy as np


def _ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def _call1(L, name, x):
    y = np.empty_like(x)
    getattr(L, name)(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    return y


def test_sincos(oracle):
    L = oracle.lib()
    x = np.linspace(-3.5, 13.0, 400001).astype(np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    L.orc_dm_sincos(x.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    rs, rc = np.sin(x.astype(np.float64)), np.cos(x.astype(np.float64))
    assert np.max(np.abs(s - rs)) < 2.5e-7 and np.max(np.abs(c - rc)) < 2.5e-7
    # away from the zeros the relative error is a few ULP
    m = np.abs(rs) > 0.05
    assert _ulp_diff(s[m], rs[m].astype(np.float32)).max() <= 4
    m = np.abs(rc) > 0.05
    assert _ulp_diff(c[m], rc[m].astype(np.float32)).max() <= 4


def test_acos(oracle):
    x = np.linspace(-1.0, 1.0, 400001).astype(np.float32)
    y = _call1(oracle.lib(), "orc_dm_acos", x)
    ref = np.arccos(x.astype(np.float64))
    assert np.max(np.abs(y - ref)) < 6e-7
    assert y[0] == np.float32(np.pi) and y[-1] == 0.0


def test_atan2(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    a = rng.normal(size=300000).astype(np.float32)
    b = rng.normal(size=300000).astype(np.float32)
    a[:4] = [0.0, 0.0, 1.0, -1.0]
    b[:4] = [1.0, -1.0, 0.0, 0.0]
    y = np.empty_like(a)
    L.orc_dm_atan2(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_size_t(a.size))
    ref = np.arctan2(a.astype(np.float64), b.astype(np.float64))
    assert np.max(np.abs(y - ref)) < 6e-7
    assert (np.abs(y) <= np.float32(np.pi)).all()


def test_exp(oracle):
    x = np.linspace(-20.0, 20.0, 200001).astype(np.float32)
    y = _call1(oracle.lib(), "orc_dm_exp", x)
    ref = np.exp(x.astype(np.float64))
    assert np.max(np.abs(y - ref) / ref) < 3e-7


def test_polar_code_round_trip(oracle):
    """encodeNormal/decodeNormal (common_device.cuh:35-65) built on detmath: the 16-bit polar code of a
    decoded direction is a fixed point (what makes G-buffer normals stable across frames)."""
    L = oracle.lib()
    q_phi = np.arange(0, 65536, 97, dtype=np.uint32)
    q_theta = np.arange(1, 65535, 89, dtype=np.uint32)
    qp, qt = np.meshgrid(q_phi, q_theta)
    phi = (np.float32(2 * np.pi) * (qp.astype(np.float32) / np.float32(65535.0))).astype(np.float32).ravel()
    theta = (np.float32(np.pi) * (qt.astype(np.float32) / np.float32(65535.0))).astype(np.float32).ravel()
    sp, cp, st, ct = (np.empty_like(phi) for _ in range(4))
    L.orc_dm_sincos(phi.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p), cp.ctypes.data_as(C.c_void_p), C.c_size_t(phi.size))
    L.orc_dm_sincos(theta.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), ct.ctypes.data_as(C.c_void_p), C.c_size_t(phi.size))
    v = np.stack([-sp * st, ct, cp * st], axis=1)
    assert np.abs(np.linalg.norm(v.astype(np.float64), axis=1) - 1.0).max() < 1e-6
    th2 = _call1(L, "orc_dm_acos", np.clip(v[:, 1], -1, 1).astype(np.float32))
    # acos is ill-conditioned at the poles in fp32 (true of libm too): one 16-bit step there, a few
    # ULP elsewhere
    mid = np.abs(theta - np.float32(np.pi / 2)) < 1.3
    assert np.abs(th2 - theta)[mid].max() < 4e-6
    assert np.abs(th2 - theta).max() <= np.float32(np.pi) / 65535 * 1.01
