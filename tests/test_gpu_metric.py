"""metrics.rs on the GPU: reference KATs (metrics_tests.rs:4-33) and bit-exactness against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float32).eps


@pytest.fixture(scope="module")
def r():
    import redis_hnsw_b200 as r

    return r


def test_reference_kats(r):
    def one(a, b):
        return float(r.l2_batch(a[None, :], b[None, :])[0])

    assert abs(one(np.full(512, 1.0, np.float32), np.full(512, 1.0, np.float32)) - 0.0) < EPS      # :4-9
    assert abs(one(np.zeros(512, np.float32), np.ones(512, np.float32)) - -512.0) < EPS              # :12-17
    assert abs(one(np.zeros(512, np.float32), np.full(512, 512.0, np.float32)) - -134217728.0) < EPS  # :20-25
    assert abs(one(np.zeros(33, np.float32), np.ones(33, np.float32)) - -33.0) < EPS                # :28-33


@pytest.mark.parametrize("dim", [4, 20, 33, 100, 32, 64, 96, 128, 512, 768, 1024])
def test_bit_exact_vs_oracle(r, oracle_mod, dim):
    rng = np.random.default_rng(dim)
    n = 20000
    a = rng.standard_normal((n, dim)).astype(np.float32)
    b = rng.standard_normal((n, dim)).astype(np.float32)
    a[:50] = b[:50]  # identical rows -> -0.0
    got = r.l2_batch(a, b)
    want = oracle_mod.euclidean_batch(a, b)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.all(np.signbit(got[:50])) and np.all(got[:50] == 0.0)


def test_empty_and_errors(r):
    assert r.l2_batch(np.zeros((0, 32), np.float32), np.zeros((0, 32), np.float32)).shape == (0,)
