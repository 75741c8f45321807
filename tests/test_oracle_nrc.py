"""The NRC oracle (oracle/nrc.cpp) pinned without a GPU: layer sizes of the tcnn config, encoding
properties, and that Adam+EMA on the RelativeL2Luminance loss actually learns a target."""
import numpy as np

from gfxexp_b200 import engine


def _queries(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.02, 0.98, size=(n, 14)).astype(np.float32)


def test_parameter_count_matches_tcnn_config(oracle):
    net = oracle.OracleNrc(2, 1e-2)
    # 64x64 + 64x64 + 16x64 matrix weights; level 0 dense 16^3 = 4096, levels 1..15 2^15 entries, F = 2
    assert net.num_matrix_weights == 64 * 64 * 2 + 16 * 64
    assert net.num_params == net.num_matrix_weights + (4096 + 15 * 32768) * 2  # 991 232 grid halfs
    net5 = oracle.OracleNrc(5, 1e-2)
    assert net5.num_matrix_weights == 64 * 64 * 5 + 16 * 64


def test_encoding_properties(oracle):
    net = oracle.OracleNrc(2, 1e-2)
    net.set_params(engine.random_nrc_params(net.num_params, net.num_matrix_weights, grid_amplitude=0.5))
    q = _queries(512)
    e = net.encode(q).astype(np.float32)
    # one-blob: 4 bins per dim are a partition of unity (wrap-around CDF differences)
    ob = e[:, 32:52].reshape(-1, 5, 4)
    assert np.abs(ob.sum(axis=2) - 1.0).max() < 4e-3
    assert (ob >= -1e-3).all()
    # identity + padding
    assert np.array_equal(e[:, 52:58], q[:, 8:14].astype(np.float16).astype(np.float32))
    assert (e[:, 58:] == 1.0).all()
    # hash grid is continuous: a tiny move in position changes the coarse levels (res <= 2048) only slightly
    q2 = q.copy()
    q2[:, :3] += 1e-5
    e2 = net.encode(q2).astype(np.float32)
    assert np.abs(e2[:, :16] - e[:, :16]).max() < 0.06


def test_training_learns_target(oracle):
    net = oracle.OracleNrc(2, 1e-2)
    net.set_params(engine.random_nrc_params(net.num_params, net.num_matrix_weights))
    q = _queries(2048, 3)
    target = np.stack([0.5 + 0.4 * np.sin(6 * q[:, 0]), q[:, 1] * q[:, 8], 0.3 + 0.5 * q[:, 2]], axis=1).astype(np.float32)
    losses = [net.train(q, target) for _ in range(40)]
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.25 * losses[0], losses[::8]
    pred = net.infer(q)
    assert np.mean((pred - target) ** 2) < 0.05


def test_one_blob_follows_the_soa_kernel(oracle):
    """The reference's composite encoding runs OneBlob on an SoA slice, i.e. tiny-cuda-nn's kernel_one_blob_soa
    (oneblob.h:110-139): features 32..51 of the oracle's encoding equal a float32 numpy transcription of that kernel bit for
    bit (numpy float32 arithmetic is IEEE, one rounding per operation, like the oracle's)."""
    f = np.float32
    rng = np.random.default_rng(3)
    n = 20000
    q = rng.uniform(0.0, 1.0, size=(n, 14)).astype(f)
    q[:200, 3:8] = rng.choice(np.array([0.0, 0.25, 0.5, 0.75, 1.0], dtype=f), size=(200, 5))
    net = oracle.OracleNrc(2, 1e-2)
    net.set_params(np.zeros(net.num_params, dtype=np.float16))
    got = net.encode(q).view(np.uint16).reshape(n, 64)[:, 32:52]

    def qc(x):
        u = (x * f(4.0)).astype(f)
        u2 = (u * u).astype(f)
        u4 = (u2 * u2).astype(f)
        inner = ((f(1.0) - (f(f(2.0) / f(3.0)) * u2).astype(f)).astype(f) + (f(f(1.0) / f(5.0)) * u4).astype(f)).astype(f)
        v = (((f(f(15.0) / f(16.0)) * u).astype(f) * inner).astype(f) + f(0.5)).astype(f)
        return np.maximum(f(0.0), np.minimum(f(1.0), v)).astype(f)

    def cdf3(t):
        return ((qc(t) + qc((t - f(1.0)).astype(f))).astype(f) + qc((t + f(1.0)).astype(f))).astype(f)

    want = np.empty((n, 20), dtype=np.uint16)
    for d in range(5):
        x = q[:, 3 + d]
        left = cdf3((-x).astype(f))
        for k in range(4):
            right = cdf3((f(0.25 * (k + 1)) - x).astype(f))
            want[:, d * 4 + k] = (right - left).astype(f).astype(np.float16).view(np.uint16)
            left = right
    assert np.array_equal(got, want), int((got != want).sum())
