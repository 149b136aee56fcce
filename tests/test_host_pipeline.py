"""Host-side product logic (no GPU): reader + normaliser + tiler + sparse post-processing against reference dumps."""
import hashlib
import os

import numpy as np
import pytest

from mustache_b200 import fdr, ladder, normalize, postprocess, readers, tiler
from oracle import postprocess as opost
from oracle import scalespace as osc
from tests import synth

G = synth.GOLDEN


@pytest.fixture(scope="module")
def chr21(tmp_path_factory):
    d = tmp_path_factory.mktemp("chr21")
    raw, kr = synth.write_chr21_text(str(d))
    x, y, v = readers.read_text(raw, 2000000, kr, "21", 5000)
    dpx = tiler.distance_in_px(2000000, 5000)
    normalize.normalize_sparse(x, y, v, 5000, dpx)
    return x, y, v, dpx


def test_reader_and_normaliser_bit_exact(chr21):
    x, y, v, dpx = chr21
    z = np.load(os.path.join(G, "chr21_blocks.npz"))
    assert dpx == int(z["dpx"]) and len(v) == int(z["nnz"])
    dig = hashlib.sha256(np.ascontiguousarray(x, np.int64).tobytes() + np.ascontiguousarray(y, np.int64).tobytes()
                         + np.ascontiguousarray(v, np.float64).tobytes()).digest()
    assert np.array_equal(np.frombuffer(dig, np.uint8), z["coo_digest"])


def test_block_geometry(chr21):
    x, y, v, dpx = chr21
    z = np.load(os.path.join(G, "chr21_blocks.npz"))
    n = int(max(x.max(), y.max()) + 1)
    chunk, starts, ends = tiler.block_geometry(n, dpx)
    assert chunk == 2000 and starts == list(z["start"]) and ends == list(z["end"])
    assert tiler.block_geometry(1500, 400) == (2000, [0], [1500])
    assert [tiler.block_mask_size(i, starts, ends, dpx) for i in range(len(starts))] == [-1, 400, 400, 400, 400, 779]


@pytest.mark.parametrize("b", [1, 2, 3, 4, 5])
def test_sparse_postprocess_matches_reference_loops(chr21, b):
    """Feed the reference's own (row, col, p_raw, scale) dump of each chr21 block to the product's sparse
    post-processing; the loops must equal what mustache() returned for that block (coordinates, FDR, scale)."""
    x, y, v, dpx = chr21
    z = np.load(os.path.join(G, "chr21_blocks.npz"))
    start, end = int(z["start"][b]), int(z["end"][b])
    xc, yc, vc = tiler.block_coo(x, y, v, start, end)
    mr, mc, mv = tiler.block_mask_pixels(xc, yc, vc, 2000)
    assert len(mr) == int(z["b%d_nz_count" % b])
    loops, _ = postprocess.call_loops(2000, dpx, start, mr, mc, mv, z["b%d_rows" % b], z["b%d_cols" % b],
                                      z["b%d_p" % b], z["b%d_scale" % b], st=0.8, pt=0.1)
    ref = z["b%d_loops" % b]
    got = np.array(loops, float).reshape(-1, 4)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name", list(synth.SYNTH_TILES))
def test_sparse_postprocess_synthetic(name):
    spec = synth.SYNTH_TILES[name]
    z = np.load(os.path.join(G, "synth_%s.npz" % name))
    c = synth.make_tile(**spec["gen"])
    n = c.shape[0]
    r, cc = np.nonzero((c != 0) & (np.subtract.outer(np.arange(n), np.arange(n)) <= -4))
    loops, _ = postprocess.call_loops(n, spec["dpx"], 0, r, cc, c[r, cc], z["rows"], z["cols"], z["p"], z["scale"],
                                      st=spec["st"], pt=spec["pt"])
    assert np.array_equal(np.array(loops, float).reshape(-1, 4), z["loops"])


def test_small_mask_returns_nothing():
    r = np.arange(100)
    loops, _ = postprocess.call_loops(2000, 400, 0, r, r + 10, np.ones(100), r[:5], r[:5] + 10, np.full(5, 1e-9),
                                      np.full(5, 2.1), st=0.0, pt=0.5)
    assert loops == []


def test_fdr_matches_statsmodels_form():
    rng = np.random.default_rng(3)
    p = rng.random(4000) ** 4
    p[10:20] = p[10]
    assert np.array_equal(fdr.fdr_bh(p), opost.bh_statsmodels_form(p))
    assert fdr.fdr_bh(np.zeros(0)).size == 0


def test_program_matches_oracle_ladder():
    for octs in ([1.6, 3.2], [1.6, 3.2, 6.4, 12.8], [1.6], [2.0, 4.0, 8.0]):
        lad = osc.sigma_ladder(octs)
        prog = ladder.build_program(octs, dedupe=False)
        assert len(prog.steps) == len(lad)
        for st, lv in zip(prog.steps, lad):
            assert st.sigma == lv["sigma"] and st.radius == lv["radius"] and np.array_equal(st.taps, lv["taps"])
            assert st.restart == (lv["k"] == 1)
            i = lv["k"] - 1
            assert st.score_id == (lv["octave"] * 12 + i if 3 <= i <= 11 else 0)
            if st.score_id:
                assert st.score_sigma == osc.level_sigma(octs[lv["octave"]], i)
        ded = ladder.build_program(octs, dedupe=True)
        assert len(ded.steps) == len(lad) - 2 * (len(octs) - 1)        # default-style ladders share 2 levels/octave
        assert [s.score_id for s in ded.steps if s.score_id] == [s.score_id for s in prog.steps if s.score_id]
    odd = ladder.build_program([1.6, 3.0], dedupe=True)                 # not a factor of 2: nothing to share
    assert len(odd.steps) == 24 and odd.steps[12].restart


def test_mask_pixels_last_write_wins():
    xc = np.array([3, 3, 5, 0]); yc = np.array([9, 9, 8, 2]); vc = np.array([1.0, 7.0, 2.0, 4.0])
    r, c, v = tiler.block_mask_pixels(xc, yc, vc, 100)
    assert list(r) == [3] and list(c) == [9] and list(v) == [7.0]       # (5,8): d=3 < 4; (0,2): d=2


@pytest.mark.parametrize("seed,n,dpx,missing,st,pt", [(101, 240, 90, 0.05, 0.6, 0.3), (102, 260, 300, 0.25, 0.4, 0.5),
                                                      (103, 224, 70, 0.0, 0.88, 0.2), (104, 300, 120, 0.4, 0.3, 0.8)])
def test_sparse_postprocess_equals_dense_oracle(seed, n, dpx, missing, st, pt):
    """Seeded tiles the reference never saw (sparse masks, band wider than the tile, loose and tight thresholds): the
    product's sparse post-processing must return exactly what the dense restatement of mustache.py:774-850 returns
    from the same per-pixel state (coordinates, FDR and scale bit for bit; candidates near the tile border exercise the
    negative-slice quirk of the sparsity windows)."""
    c = synth.make_tile(n=n, dpx=dpx, seed=seed, blob_seed=seed + 50, nblobs=14, missing=missing)
    res = osc.scale_space(c, dpx, [1.6, 3.2], use_scipy=True)
    assert not res["skipped"] and res["nz_count"] >= 10000
    nz, filled = osc.mask_and_fill(c, dpx)
    ref = opost.loops_dense(filled, nz, res["p"], res["scale"], 7, dpx, st, pt)
    found = res["p"] != 2
    mr, mc = res["rows"], res["cols"]
    loops, _ = postprocess.call_loops(n, dpx, 7, mr, mc, c[mr, mc], mr[found], mc[found], res["p"][found],
                                      res["scale"][found], st=st, pt=pt)
    key = lambda l: (l[0], l[1])
    assert len(ref) > 0 and sorted(map(tuple, loops), key=key) == sorted(map(tuple, ref), key=key)


def test_normaliser_equals_literal_restatement(chr21):
    """mustache_b200.normalize selects each diagonal's contacts from one stable sort; the literal restatement in
    oracle/normalize.py builds `distances == d` per diagonal like mustache.py:632-633.  Same bits, both branches,
    unsorted input, empty diagonals, contacts beyond the visited diagonals."""
    from oracle import normalize as onorm
    rng = np.random.default_rng(17)
    n, dpx = 3000, 400
    x = rng.integers(0, n - 450, size=9000)
    d = rng.integers(0, dpx + 30, size=9000)
    d[d == 33] = 34                                   # an empty diagonal
    y = x + d
    key = np.unique(x * 100000 + y, return_index=True)[1]
    key = key[rng.permutation(len(key))]              # unsorted
    x, y = x[key], y[key]
    v = rng.gamma(2.0, 3.0, size=len(x))
    for res in (5000, 500):                           # windowed branch / global branch ((n - dpx) * res <= 2 Mb)
        a, b = v.copy(), v.copy()
        wa = normalize.normalize_sparse(x, y, a, res, dpx)
        wb = onorm.normalize_sparse(x, y, b, res, dpx)
        assert wa == wb and np.array_equal(a, b)
        assert not np.array_equal(a, v)


def test_integer_counts_keep_reference_truncation(tmp_path):
    """read_pd without a bias file on integer text yields int64 values and normalize_sparse truncates its z-scores into
    them (mustache.py:668, 683): the reader keeps the dtype, so the CLI pipeline reproduces what the literal restatement
    of the reference's normaliser (oracle/normalize.py) does on the same integer array."""
    from mustache_b200 import normalize, readers
    from oracle import normalize as onorm
    p = str(tmp_path / "ints.txt")
    rng = np.random.default_rng(3)
    with open(p, "w") as f:
        for i in range(300):
            for d in range(0, 30):
                if i + d < 300:
                    f.write("chr1\t%d\tchr1\t%d\t%d\n" % (i * 5000, (i + d) * 5000, rng.integers(1, 40)))
    x, y, v = readers.read_text(p, 2000000, False, "chr1", 5000)
    assert v.dtype.kind == "i"
    ref = v.copy()
    onorm.normalize_sparse(np.asarray(x), np.asarray(y), ref, 5000, 400)
    normalize.normalize(x, y, v, 5000, 400, eng=None)
    assert v.dtype.kind == "i" and np.array_equal(v, ref) and set(np.unique(v)) <= {-2, -1, 0, 1, 2} and (v != 0).any()


def test_block_slicer_equals_block_coo():
    """tiler.BlockSlicer (one sort, contiguous row ranges) selects exactly what the reference's boolean masks select
    (mustache.py:919-922), in the same order, also for unsorted input with duplicate coordinates."""
    from mustache_b200 import tiler
    rng = np.random.default_rng(11)
    n = 700
    x = rng.integers(0, n, 5000)
    y = np.minimum(x + rng.integers(0, 60, 5000), n - 1)
    v = rng.random(5000)
    x[100:110], y[100:110] = x[90:100], y[90:100]                 # duplicates: last write wins
    sl = tiler.BlockSlicer(x, y, v)
    chunk, starts, ends = 260, [0, 130, 440], [260, 390, 700]
    for s0, e0 in zip(starts, ends):
        a = tiler.block_mask_pixels(*tiler.block_coo(x, y, v, s0, e0), chunk)
        b = tiler.block_mask_pixels(*sl.block(s0, e0), chunk)
        for p, q in zip(a, b):
            assert np.array_equal(p, q)
    xs = np.sort(x)
    sl2 = tiler.BlockSlicer(xs, y, v)                             # already row sorted: no copy, same answer as the masks
    a = tiler.block_coo(xs, y, v, 130, 390)
    b = sl2.block(130, 390)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))


@pytest.mark.parametrize("seed,n,dpx,missing,st,pt", [(101, 240, 90, 0.05, 0.6, 0.3), (102, 260, 300, 0.25, 0.4, 0.5),
                                                       (103, 300, 120, 0.4, 0.2, 0.8)])
def test_candidate_postprocess_equals_record_postprocess(seed, n, dpx, missing, st, pt):
    """The split the product uses (device: BH, o < pt, sparsity, 3 x 3 neighbourhoods; host: enrichment + clustering from
    those candidates, postprocess.call_loops_from_candidates) gives the loops of the all-host path on the full record
    list (postprocess.call_loops, itself pinned to the dense oracle and the reference dumps above).  The device half is
    stood in for by its dense restatement on the oracle (tests/test_sharded_blocks.py:OracleEngine.candidates_batch)."""
    from tests.test_sharded_blocks import OracleEngine
    c = synth.make_tile(n=n, dpx=dpx, seed=seed, blob_seed=seed + 50, nblobs=14, missing=missing)
    r, cc = np.nonzero(np.triu(c, 4))
    old = postprocess.MIN_MASK_FOR_BH
    postprocess.MIN_MASK_FOR_BH = 100
    try:
        eng = OracleEngine()
        eng.set_octaves([1.6, 3.2])
        eng.configure(n, dpx, 1)
        eng.upload_coo(0, r, cc, c[r, cc])
        eng.run()
        rec = eng.records_batch()[0]
        eng.select_candidates(pt, st)
        cand = eng.candidates_batch()[0]
        a, _ = postprocess.call_loops(n, dpx, 7, r, cc, c[r, cc], rec["rows"], rec["cols"], rec["p"], rec["sigma"], st, pt)
        b = postprocess.call_loops_from_candidates(n, dpx, 7, r, cc, c[r, cc], cand)
    finally:
        postprocess.MIN_MASK_FOR_BH = old
    assert len(a) > 0 and a == b


def test_native_parser_equals_pandas(tmp_path):
    """mb200_contacts_open (native multi-threaded parser) returns exactly what the pandas path of read_pd returns: same rows,
    same order, same values, same inferred dtype -- on the bundled chr21 file, a multi-chromosome integer file (> 1 MB, so
    that several threads split it), a 3-column file and CRLF input; and hands anything unusual back to pandas."""
    from mustache_b200 import readers
    raw, kr = synth.write_chr21_text(str(tmp_path))

    def both(path, chrom):
        a = readers.parse_contacts_native(path, chrom)
        b = readers._parse_contacts_pandas(path, chrom)
        return a, b
    a, b = both(raw, "21")
    assert a is not None and len(a[0]) == 677085 and a[3] == 5
    for p, q in zip(a[:3], b[:3]):
        assert p.dtype == q.dtype and np.array_equal(p, q)
    rng = np.random.default_rng(0)
    multi = str(tmp_path / "multi.txt")
    with open(multi, "w") as f:
        for i in range(90000):
            c = "chr%d" % (1 + i % 3) if i % 2 else str(1 + i % 3)          # 'chr2' and '2' name the same chromosome
            f.write("%s\t%d\t%s\t%d\t%d\n" % (c, 5000 * rng.integers(0, 4000), c, 5000 * rng.integers(0, 4000), rng.integers(1, 300)))
    assert os.path.getsize(multi) > (1 << 20)
    a, b = both(multi, "chr2")
    assert a[2].dtype == np.int64 and len(a[0]) == 30000
    for p, q in zip(a[:3], b[:3]):
        assert p.dtype == q.dtype and np.array_equal(p, q)
    three = str(tmp_path / "three.txt")
    vals = ["0.125", "12.345678", "1e-3", "7", "3.0", "123456789012345", "0.000123456789", "2.5E2", "-4.75"]
    with open(three, "w") as f:
        for k, v in enumerate(vals):
            f.write("%d %d %s\r\n" % (1000 * k, 1000 * k + 5000, v))
    a, b = both(three, "1")
    assert a[3] == 3 and a[2].dtype == np.float64
    for p, q in zip(a[:3], b[:3]):
        assert np.array_equal(p, q)
    for bad in ("chr1\t100\tchr1\t200\n", "chr1\t100\tchr1\t200\tNaN\n", "chr1\t100.5\tchr1\t200\t3\n",
                "chr1\t100\tchr1\t200\t0.12345678901234567\n", "\"chr1\"\t100\tchr1\t200\t3\n"):
        p = str(tmp_path / "bad.txt")
        open(p, "w").write("chr1\t0\tchr1\t5000\t2\n" + bad)
        assert readers.parse_contacts_native(p, "chr1") is None
    # the whole reader on both parsers
    os.environ["MUSTACHE_READER"] = "pandas"
    try:
        ref = readers.read_text(raw, 2000000, kr, "21", 5000)
    finally:
        del os.environ["MUSTACHE_READER"]
    got = readers.read_text(raw, 2000000, kr, "21", 5000)
    assert all(np.array_equal(p, q) for p, q in zip(got, ref))
