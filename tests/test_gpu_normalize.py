"""GPU parity of the device normaliser (mb200_normalize_sparse) against the numpy restatement of the reference's
normalize_sparse (mustache.py:622-686), itself identical to the reference call for call.

Tolerances: np.mean / np.std are reproduced with numpy's pairwise summation, so the global branch and the per-diagonal
weights must be bit-exact; the 2 Mb box sums run left to right instead of through the CPU-dependent BLAS dot product of
np.convolve, so windowed z-scores are compared to 1e-10 relative (observed ~1e-13)."""
import ctypes as C
import math

import numpy as np
import pytest

from mustache_b200 import mustache as mm
from mustache_b200 import normalize, synth as gen, tiler
from tests import synth as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return mm.get_engine()


@pytest.fixture(scope="module")
def chr21_coo(tmp_path_factory):
    """chr21 contacts exactly as the text reader hands them to normalize_sparse."""
    from mustache_b200 import readers
    d = tmp_path_factory.mktemp("chr21n")
    raw, kr = fx.write_chr21_text(str(d))
    x, y, v = readers.read_text(raw, 2000000, kr, "21", 5000)
    return np.asarray(x), np.asarray(y), np.asarray(v, dtype=np.float64), tiler.distance_in_px(2000000, 5000)


def test_windowed_branch_chr21(eng, chr21_coo):
    x, y, v, dpx = chr21_coo
    ref = v.copy()
    w_ref = normalize.normalize_sparse(x, y, ref, 5000, dpx)
    got = v.copy()
    w_got = normalize.normalize_sparse_device(eng, x, y, got, 5000, dpx)
    assert len(w_got) == len(w_ref) == dpx + 2
    assert w_got == w_ref                                       # per-diagonal np.mean: bit-exact
    assert np.all(np.isfinite(got))
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    assert err.max() < 1e-10, err.max()


def test_windowed_branch_edges(eng):
    """Sparse synthetic chromosome: windows with < 30 contacts (global fallback), empty diagonals, single-contact
    diagonals (std 0 -> division by zero -> 0), contacts beyond dpx + 1 (left untouched)."""
    rng = np.random.default_rng(7)
    n, res, dpx = 3000, 5000, 400
    x = rng.integers(0, n - 500, size=6000)
    d = rng.integers(0, dpx + 40, size=6000)
    d[d == 17] = 18                                             # diagonal 17 stays empty
    x = np.concatenate([x, [5]]); d = np.concatenate([d, [17 + 600]])
    y = x + d
    key = np.unique(x * 100000 + y, return_index=True)[1]
    x, y = x[key], y[key]
    v = rng.gamma(2.0, 3.0, size=len(x))
    ref = v.copy()
    w_ref = normalize.normalize_sparse(x, y, ref, res, dpx)
    got = v.copy()
    w_got = normalize.normalize_sparse_device(eng, x, y, got, res, dpx)
    assert w_got == w_ref
    far = np.abs(y - x) >= dpx + 2
    assert far.any() and np.array_equal(got[far], v[far])       # the diagonal loop never visits them
    assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())


def test_global_branch_bit_exact(eng):
    """(n - dpx) * res <= 2 Mb: plain per-diagonal z-score with np.mean / np.std -> bit-exact."""
    n, res = 350, 5000
    dpx = tiler.distance_in_px(2000000, res)
    x, y, v = gen.poisson_chromosome(n, dpx, lam_scale=200.0, seed=99, nloops=10, loop_boost=30.0)
    v = v.astype(np.float64)
    v[3] = np.nan                                               # cleaned by the leading nan_to_num (mustache.py:672)
    ref = v.copy()
    normalize.normalize_sparse(x, y, ref, res, dpx)
    got = v.copy()
    w = normalize.normalize_sparse_device(eng, x, y, got, res, dpx)
    assert w == []
    assert np.array_equal(got, ref)


def test_pairwise_mean_std_large_diagonals(eng):
    """Diagonals long enough for several levels of numpy's pairwise recursion (> 128 * 8 elements)."""
    rng = np.random.default_rng(5)
    n, res, dpx = 60000, 5000, 6
    xs, ys = [], []
    for d in range(dpx):
        xi = np.sort(rng.choice(n - d, size=int((n - d) * 0.7), replace=False))
        xs.append(xi); ys.append(xi + d)
    x, y = np.concatenate(xs), np.concatenate(ys)
    v = rng.lognormal(0.0, 1.0, size=len(x))
    # global branch forced by a huge distance
    ref = v.copy(); normalize.normalize_sparse(x, y, ref, 10, n + 5)
    got = v.copy(); normalize.normalize_sparse_device(eng, x, y, got, 10, n + 5)
    assert np.array_equal(got, ref)


def test_cli_chr21_with_device_normaliser(tmp_path, monkeypatch):
    """README command with MUSTACHE_NORMALIZE=device: same 90 loops, scales exact, FDR within the north_star tolerance."""
    from tests.test_gpu_e2e import G, _same_tsv
    import os
    monkeypatch.setenv("MUSTACHE_NORMALIZE", "device")
    raw, kr = fx.write_chr21_text(str(tmp_path))
    out = str(tmp_path / "chr21_out.tsv")
    mm.main(["-f", raw, "-b", kr, "-ch", "21", "-r", "5kb", "-pt", "0.1", "-st", "0.8", "-o", out])
    assert _same_tsv(out, os.path.join(G, "chr21_loops.tsv")) == 90


def test_windowed_branch_other_window(eng):
    """2 kb bins: window of 1000 bins, distance limit 1000 bins, both window ends clipped for most contacts."""
    rng = np.random.default_rng(11)
    n, res, dpx = 2500, 2000, 1000
    m = 120000
    x = rng.integers(0, n, size=m)
    d = np.minimum(rng.geometric(0.004, size=m) - 1, dpx + 1)
    y = x + d
    ok = y < n
    x, y = x[ok], y[ok]
    key = np.unique(x * 100000 + y, return_index=True)[1]
    x, y = x[key], y[key]
    v = rng.gamma(1.5, 2.0, size=len(x)) / (1.0 + 0.01 * (y - x))
    ref = v.copy()
    w_ref = normalize.normalize_sparse(x, y, ref, res, dpx)
    got = v.copy()
    w_got = normalize.normalize_sparse_device(eng, x, y, got, res, dpx)
    assert w_got == w_ref and len(w_ref) == dpx + 2
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_unsorted_input(eng):
    """Contacts in random order: the device groups them by (diagonal, position) itself.  np.mean then sees another order
    than numpy's, so everything is compared at tolerance level."""
    rng = np.random.default_rng(23)
    n, res, dpx = 3000, 5000, 400
    x, y, c = gen.synthetic_chromosome(n, dpx, 18.0, seed=5, nloops=20, loop_dmax=30)
    v = c.astype(np.float64) * rng.uniform(0.7, 1.3, size=len(c))
    perm = rng.permutation(len(v))
    x, y, v = x[perm], y[perm], v[perm]
    ref = v.copy()
    w_ref = normalize.normalize_sparse(x, y, ref, res, dpx)
    got = v.copy()
    w_got = normalize.normalize_sparse_device(eng, x, y, got, res, dpx)
    assert np.allclose(w_got, w_ref, rtol=1e-13, atol=0)
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_config3_input_at_1kb(eng):
    """BASELINE configs[2] input (1.3 M contacts, 2 002 diagonals, 2 000-bin windows) with a non-trivial bias so that no
    window holds identical values: device vs numpy normaliser."""
    spec = dict(gen.CONFIG3)
    res = spec.pop("res")
    spec["n"] = 12000                                           # 5 blocks' worth: keeps the numpy side at ~15 s
    x, y, c = gen.synthetic_chromosome(**spec)
    bias = np.random.default_rng(5).uniform(0.6, 1.6, size=spec["n"])
    v = c / bias[x] / bias[y]
    ref = v.copy()
    w_ref = normalize.normalize_sparse(x, y, ref, res, spec["dpx"])
    got = v.copy()
    w_got = normalize.normalize_sparse_device(eng, x, y, got, res, spec["dpx"])
    assert w_got == w_ref
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
