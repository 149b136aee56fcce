"""GPU parity: edge cases and size-independent properties (empty / tiny masks, whole-chromosome tiling on synthetic
counts incl. the n <= CHUNK case and the right-aligned last block, crop invariance at the BASELINE tile size)."""
import numpy as np
import pytest

from mustache_b200 import mustache as mm
from mustache_b200 import normalize, synth as gen, tiler
from oracle import postprocess as opost
from oracle import scalespace as osc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return mm.get_engine()


def test_empty_and_tiny_masks(eng):
    eng.set_octaves([1.6, 3.2])
    mm._PROGRAM_KEY[id(eng)] = (1.6, 3.2)
    eng.configure(256, 100, 3)
    z = np.zeros((256, 256))
    few = z.copy()
    few[np.arange(40), np.arange(40) + 9] = 1.5            # 40 mask pixels: below the 50 of mustache.py:701
    low = z.copy()
    low[np.arange(200), np.arange(200) + 3] = 1.0          # only diagonal 3: not in the mask at all
    for b, t in enumerate((z, few, low)):
        eng.upload_dense(b, t)
    eng.run()
    assert [eng.counts(b)[0] for b in range(3)] == [0, 40, 0]
    assert eng.records(0)["n_found"] == 0 and eng.records(2)["n_found"] == 0
    # the drop-in functions return [] exactly where the reference does
    assert mm.mustache(few.copy(), "1", "1", 5000, [], 0, 256, -1, 100, [1.6, 3.2], 0.8, 0.1) == []
    assert mm.mustache(z.copy(), "1", "1", 5000, [], 0, 256, -1, 100, [1.6, 3.2], 0.8, 0.1) == []


def _oracle_blocks(x, y, v, n, dpx, octaves, st, pt):
    """The reference's regulator()/process_block() flow (mustache.py:896-960) on the oracle."""
    chunk, start, end = tiler.block_geometry(n, dpx)
    out = []
    for b in range(len(start)):
        xc, yc, vc = tiler.block_coo(x, y, v, start[b], end[b])
        cc = np.zeros((chunk, chunk))
        cc[xc, yc] = vc
        res = osc.scale_space(cc, dpx, octaves, use_scipy=True)
        loops = []
        if not res["skipped"]:
            nz, filled = osc.mask_and_fill(cc, dpx)
            loops = opost.loops_dense(filled, nz, res["p"], res["scale"], start[b], dpx, st, pt)
        ms = tiler.block_mask_size(b, start, end, dpx)
        out += [l for l in loops if l[0] >= start[b] + ms or l[1] >= start[b] + ms]
    return out


@pytest.mark.parametrize("n,res,dist_bp", [(1500, 5000, 2000000), (3900, 10000, 2000000)])
def test_whole_chromosome_matches_oracle(n, res, dist_bp):
    """Synthetic Poisson chromosome through normalise -> tile -> engine -> post-process vs the same flow on the oracle:
    n <= CHUNK (one zero-padded block, global z-score branch of normalize_sparse) and a 3-block chromosome whose last
    block is right-aligned and overlaps the previous one by 1 900 bins (mustache.py:909-910)."""
    dpx = tiler.distance_in_px(dist_bp, res)
    x, y, v = gen.poisson_chromosome(n, dpx, lam_scale=200.0, seed=3000 + n, nloops=60, loop_boost=30.0)
    normalize.normalize_sparse(x, y, v, res, dpx)
    got = mm.call_blocks(x, y, v, n, dpx, [1.6, 3.2], 0.5, 0.2, verbose=False)
    ref = _oracle_blocks(x, y, v, n, dpx, [1.6, 3.2], 0.5, 0.2)
    key = lambda l: (int(l[0]), int(l[1]))
    got, ref = sorted(got, key=key), sorted(ref, key=key)
    assert len(ref) > 0 and [key(l) for l in got] == [key(l) for l in ref]
    assert [l[3] for l in got] == [l[3] for l in ref]
    assert max(abs(a[2] - b[2]) for a, b in zip(got, ref)) <= 1e-6


def test_crop_invariance_at_baseline_size(eng):
    """BASELINE configs[1] (10k x 10k dense band, dpx 5000, 4 octaves).  (v, detection scale) of a pixel only depend on
    its (rmax+2)-neighbourhood, so the records of the full tile restricted to the interior of a 1500 x 1500 crop must
    equal, bit for bit, what the oracle computes on the crop alone; two runs of the full tile must be identical."""
    n, dpx, octs = 10000, 5000, [1.6, 3.2, 6.4, 12.8]
    band = gen.dense_band_tile(n, dpx, seed=1001, blob_seed=1002, nblobs=200)
    eng.set_octaves(octs)
    mm._PROGRAM_KEY[id(eng)] = tuple(octs)
    eng.configure(n, dpx, 1)
    eng.upload_band(0, band)
    eng.run()
    full = eng.records(0)
    eng.upload_band(0, band)
    eng.run()
    again = eng.records(0)
    for k in ("rows", "cols", "v", "p", "score_id"):
        assert np.array_equal(full[k], again[k]), k                       # deterministic records
    assert full["nz_count"] == osc.contact_bins(n, dpx)
    r0, m, margin = 4000, 1500, 60
    i = np.arange(r0, r0 + m)[:, None]
    j = np.arange(r0, r0 + m)[None, :]
    d = j - i
    crop = np.zeros((m, m))
    ok = (d >= 4) & (d <= dpx + 1)
    crop[ok] = band[np.broadcast_to(i, d.shape)[ok], d[ok] - 4]
    ref = osc.scale_space(crop, dpx, octs, use_scipy=True)
    found = ref["p"] != 2
    rr, rc = ref["rows"][found], ref["cols"][found]
    inner = (rr >= margin) & (rr < m - margin) & (rc >= margin) & (rc < m - margin)
    sel = (full["rows"] >= r0 + margin) & (full["rows"] < r0 + m - margin) & (full["cols"] >= r0 + margin) & (full["cols"] < r0 + m - margin)
    assert inner.sum() > 500
    assert np.array_equal(full["rows"][sel] - r0, rr[inner]) and np.array_equal(full["cols"][sel] - r0, rc[inner])
    assert np.array_equal(full["v"][sel], ref["v"][found][inner])
    assert np.array_equal(full["sigma"][sel], ref["scale"][found][inner])
