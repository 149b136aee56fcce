"""GPU parity on the BASELINE.json configs at full size, against outputs of the UNMODIFIED reference produced in the build
container (tests/golden/make_golden.py cfg2 / cfg3 / cfg3d / cfg4 / cfg5), plus the engine paths those sizes exercise
(multi-pass batches, packed batch fetch, capacity retry, multi-pass differential).  Inputs are regenerated here from the
seeded generators in mustache_b200/synth.py; only the reference's OUTPUTS are fixtures."""
import hashlib
import os

import numpy as np
import pytest

from mustache_b200 import blockrun, normalize, synth as gen, tiler
from mustache_b200.fdr import fdr_bh
from mustache_b200 import mustache as mm
from tests import synth
from tests.test_gpu_e2e import FDR_TOL, _read_tsv

pytestmark = pytest.mark.gpu
G = synth.GOLDEN


@pytest.fixture(scope="module")
def eng():
    return mm.get_engine()


def _set(eng, octs):
    eng.set_octaves(octs)
    mm._PROGRAM_KEY[id(eng)] = tuple(float(o) for o in octs)


def _same_rows(got, ref):
    """TSV rows (already split and sorted): coordinates and DETECTION_SCALE strings exact, FDR within 1e-6."""
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g[:6] == r[:6] and g[7] == r[7]
        assert abs(float(g[6]) - float(r[6])) <= FDR_TOL


# ------------------------------------------------------------------------------------------------------------------
# configs[1]: the benchmarked 10k x 10k tile, 4 octaves -- every fit and every record pinned to the reference
# ------------------------------------------------------------------------------------------------------------------
def test_config2_tile_matches_reference_run(eng):
    z = np.load(os.path.join(G, "cfg2_tile.npz"))
    n, dpx, octs = 10000, 5000, [1.6, 3.2, 6.4, 12.8]
    band = gen.dense_band_tile(n, dpx, seed=1001, blob_seed=1002, nblobs=200)
    _set(eng, octs)
    eng.configure(n, dpx, 1)
    eng.upload_band(0, band)
    eng.run()
    rec = eng.records_batch()[0]
    assert rec["nz_count"] == int(z["nz_count"]) and rec["n_found"] == int(z["n_found"])
    h = hashlib.sha256()
    for k, dt in (("rows", np.int32), ("cols", np.int32), ("v", np.float64), ("sigma", np.float64)):
        h.update(np.ascontiguousarray(rec[k], dtype=dt).tobytes())
    assert np.array_equal(np.frombuffer(h.digest(), np.uint8), z["digest"])     # 1.34 M records: coordinates, vAll, Scales bit-exact
    step = int(z["sample_step"])
    assert np.abs(rec["p"][::step] - z["sample_p"]).max() <= 1e-12
    assert abs(rec["p"].sum() - float(z["p_sum"])) <= 1e-9 * float(z["p_sum"])
    fits = eng.fits(0)                                                          # expon.fit of the 36 scored levels (mustache.py:755)
    assert np.array_equal(fits["loc"], z["fits"][:, 0])
    assert np.abs(fits["scale"] / z["fits"][:, 1] - 1).max() <= 1e-13
    # and the loops mustache() returned for the tile: BH over 1.34 M p-values, o < pt and the sparsity filter on the
    # device (mb200_select_candidates), enrichment filter and clustering on the host from the selected candidates
    eng.select_candidates(0.1, 0.88)
    cand = eng.candidates_batch()[0]
    assert 0 < len(cand["rows"]) < rec["n_found"] // 20
    assert np.array_equal(eng.q_values(0), fdr_bh(rec["p"]))       # same IEEE operations as the host formula
    loops = mm.postprocess.call_loops_from_candidates(n, dpx, 0, *gen.band_to_coo(band, n), cand)
    got, ref = np.array(sorted(loops), float).reshape(-1, 4), z["loops"][np.lexsort((z["loops"][:, 1], z["loops"][:, 0]))]
    assert got.shape == ref.shape and np.array_equal(got[:, [0, 1, 3]], ref[:, [0, 1, 3]])
    assert np.abs(got[:, 2] - ref[:, 2]).max() <= FDR_TOL


# ------------------------------------------------------------------------------------------------------------------
# configs[2]: synthetic 50k-bin chromosome at 1 kb (N = 4000, dpx = 2000 blocks)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg3_normalised():
    spec = dict(gen.CONFIG3)
    res = spec.pop("res")
    x, y, c = gen.synthetic_chromosome(**spec)
    v = c.astype(np.float64)
    normalize.normalize_sparse(x, y, v, res, spec["dpx"])
    return x, y, v, spec["n"], spec["dpx"]


def test_config3_blocks_match_reference_dump(eng, cfg3_normalised):
    """Two of the 24 blocks (N 4000, dpx 2000, wide axis-0 tiles) against the reference's locals at mustache.py:778."""
    x, y, v, n, dpx = cfg3_normalised
    z = np.load(os.path.join(G, "cfg3_blocks.npz"))
    assert len(v) == int(z["nnz"]) and abs(v.sum() - float(z["v_sum"])) <= 1e-9 * abs(float(z["v_sum"]))
    chunk, start, end = tiler.block_geometry(n, dpx)
    assert chunk == 4000 and len(start) == 24
    _set(eng, [1.6, 3.2])
    blocks = (0, 11)
    eng.configure(chunk, dpx, len(blocks))
    slicer = tiler.BlockSlicer(x, y, v)
    for k, b in enumerate(blocks):
        eng.upload_coo(k, *tiler.block_mask_pixels(*slicer.block(start[b], end[b]), chunk))
    eng.run()
    for rec, b in zip(eng.records_batch(), blocks):
        pre = "b%d_" % b
        assert rec["nz_count"] == int(z[pre + "nz_count"])
        assert np.array_equal(rec["rows"], z[pre + "rows"]) and np.array_equal(rec["cols"], z[pre + "cols"])
        assert np.array_equal(rec["v"], z[pre + "v"]) and np.array_equal(rec["sigma"], z[pre + "scale"])
        assert np.abs(rec["p"] - z[pre + "p"]).max() <= 1e-12


def test_config3_whole_chromosome_calls(cfg3_normalised):
    """All 24 blocks through the product's block pool: the reference CLI calls no loop on this input (its sparsity filter,
    mustache.py:800-811, rejects everything at this depth) and neither may we."""
    x, y, v, n, dpx = cfg3_normalised
    ref = _read_tsv(os.path.join(G, "cfg3_loops.tsv"))
    got = mm.call_blocks(x, y, v.copy(), n, dpx, [1.6, 3.2], 0.8, 0.1, verbose=False)
    assert len(got) == len(ref) == 0


def _cli_rows(tmp_path, chroms, res, specs, extra=()):
    path = str(tmp_path / "contacts.txt")
    for k, (name, spec) in enumerate(specs):
        spec = {a: b for a, b in spec.items() if a != "res"}
        x, y, c = gen.synthetic_chromosome(**spec)
        gen.write_contact_text(path, name, x, y, c, res, mode="w" if k == 0 else "a")
    out = str(tmp_path / "out.tsv")
    mm.main(["-f", path, "-ch"] + list(chroms) + ["-r", "%dkb" % (res // 1000), "-pt", "0.1", "-st", "0.8", "-o", out] + list(extra))
    return _read_tsv(out)


def test_config3_dense_variant_cli(tmp_path):
    """Same 1 kb block geometry on a chromosome dense enough for loops to survive: product CLI vs reference CLI."""
    ref = _read_tsv(os.path.join(G, "cfg3d_loops.tsv"))
    got = _cli_rows(tmp_path, ["chrT"], 1000, [("chrT", gen.CONFIG3D)])
    assert len(ref) > 100
    _same_rows(got, ref)


# ------------------------------------------------------------------------------------------------------------------
# configs[3]: the synthetic chromosomes at 5 kb (the three shortest here; all eight in bench.py --config 4)
# ------------------------------------------------------------------------------------------------------------------
def test_config4_cli_three_chromosomes(tmp_path):
    names = ["s1", "s2", "s3"]
    ref = [r for r in _read_tsv(os.path.join(G, "cfg4_loops.tsv")) if r[0] in names]
    got = _cli_rows(tmp_path, names, 5000, [(k, gen.CONFIG4[k]) for k in names])
    key = lambda r: (r[0], int(r[1]), int(r[4]))
    assert len(ref) > 0
    _same_rows(sorted(got, key=key), sorted(ref, key=key))


# ------------------------------------------------------------------------------------------------------------------
# configs[4]: differential CLI on two synthetic 20k-bin maps
# ------------------------------------------------------------------------------------------------------------------
def test_config5_diff_cli(tmp_path):
    from mustache_b200 import diff_mustache as dm
    spec = gen.CONFIG5
    A, B = gen.config5_maps(**spec)
    fa = gen.write_contact_text(str(tmp_path / "mapA.txt"), "chrD", *A, spec["res"])
    fb = gen.write_contact_text(str(tmp_path / "mapB.txt"), "chrD", *B, spec["res"])
    out = str(tmp_path / "diff")
    dm.main(["-f1", fa, "-f2", fb, "-ch", "chrD", "-r", "5kb", "-pt", "0.05", "-pt2", "0.1", "-st", "0.8", "-o", out])
    total = 0
    for suf in ("loop1", "loop2", "diffloop1", "diffloop2"):
        ref = _read_tsv(os.path.join(G, "cfg5_%s.tsv" % suf))
        _same_rows(_read_tsv(out + "." + suf), ref)
        total += len(ref)
    assert total > 0


# ------------------------------------------------------------------------------------------------------------------
# engine paths
# ------------------------------------------------------------------------------------------------------------------
def _three_tiles():
    n, dpx = 512, 200
    return n, dpx, [gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=70 + b, blob_seed=80 + b, nblobs=12, missing=0.1), n)
                    for b in range(3)]


def _equal_records(a, b, keys=("rows", "cols", "v", "p", "score_id", "sigma")):
    for k in keys:
        assert np.array_equal(a[k], b[k]), k
    assert a["nz_count"] == b["nz_count"]


def test_multi_pass_and_batch_fetch(eng):
    """pass_blocks < nblocks (forced with mb200_set_pass_limit) must not change a bit; the packed batch fetch returns
    what the per-block fetch returns."""
    n, dpx, tiles = _three_tiles()
    _set(eng, [1.6, 3.2])
    out = {}
    for limit in (0, 1, 2):
        eng.set_pass_limit(limit)
        eng.configure(n, dpx, 3)
        for b, t in enumerate(tiles):
            eng.upload_dense(b, t)
        eng.run()
        out[limit] = (eng.records_batch(), [eng.records(b) for b in range(3)], [eng.fits(b) for b in range(3)])
    eng.set_pass_limit(0)
    eng.configure(n, dpx, 3)
    for limit in (0, 1, 2):
        batch, single, fits = out[limit]
        for b in range(3):
            assert batch[b]["n_found"] > 100
            _equal_records(batch[b], single[b])
            _equal_records(batch[b], out[0][0][b])
            assert np.array_equal(fits[b]["loc"], out[0][2][b]["loc"]) and np.array_equal(fits[b]["scale"], out[0][2][b]["scale"])


def test_capacity_retry(eng):
    """A batch whose records overflow the configured capacity is re-run with a larger one, not lost (the
    MB200_ERR_CAPACITY contract of include/mustache_b200.h)."""
    from mustache_b200.engine import EngineError
    n, dpx = 1024, 400                     # ~14 k records per tile: above the engine's minimum capacity of 4096
    tiles = [gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=70 + b, blob_seed=80 + b, nblobs=12, missing=0.1), n) for b in range(3)]
    _set(eng, [1.6, 3.2])
    masks = []
    for t in tiles:
        r, c = np.nonzero(np.triu(t, 4))
        masks.append((r, c, t[r, c]))
    eng.configure(n, dpx, 3, record_fraction=1e-4)
    for b, m in enumerate(masks):
        eng.upload_coo(b, *m)
    eng.run()
    with pytest.raises(EngineError) as err:
        eng.records_batch()
    assert err.value.code == -3
    nz, found = eng.batch_counts()
    assert (found > 4096).all()
    tasks = [blockrun.BlockTask(0, b, [m]) for b, m in enumerate(masks)]
    real_configure, calls = eng.configure, []

    def tight_first(n_, dpx_, nblocks=1, intra=True, record_fraction=-1.0):
        calls.append(record_fraction)
        return real_configure(n_, dpx_, nblocks, intra, 1e-4 if len(calls) == 1 else record_fraction)
    eng.configure = tight_first
    try:
        got = [r[0] for _, r in blockrun.run_batches(eng, tasks, n, dpx)]
    finally:
        eng.configure = real_configure
    assert len(calls) == 2 and calls[1] >= 0.25
    assert [g["n_found"] for g in got] == found.tolist()


def test_differential_multi_pass(eng):
    """npairs > pass_blocks: mb200_run_differential walks the pairs in passes with identical results."""
    from mustache_b200 import diff_mustache as dm
    spec = synth.SYNTH_DIFF
    c1, c2 = synth.make_pair(**spec["gen"])
    c3, c4 = synth.make_pair(n=256, dpx=100, seed=91)
    dm._set_octaves_diff(eng, spec["octaves"])
    res = {}
    for limit in (0, 1):
        eng.set_pass_limit(limit)
        eng.configure(256, spec["dpx"], 4)
        for b, t in enumerate((c1, c2, c3, c4)):
            eng.upload_dense(b, t)
        eng.run_differential()
        res[limit] = eng.records_batch(pair=True)
    eng.set_pass_limit(0)
    eng.configure(256, spec["dpx"], 4)
    z = np.load(os.path.join(G, "diff_synth.npz"))
    for b in range(4):
        _equal_records(res[0][b], res[1][b], keys=("rows", "cols", "v", "p", "score_id", "sigma", "pair"))
    for b, pre in ((0, "m1_"), (1, "m2_")):
        assert np.array_equal(res[1][b]["rows"], z[pre + "rows"]) and np.abs(res[1][b]["pair"] - z[pre + "pair"]).max() <= 1e-9


def test_batched_coo_upload_equals_per_block(eng):
    """mb200_upload_coo_batch (one copy per array, one scatter kernel) builds the same tiles as per-block uploads, also
    into a block range that does not start at 0 and with an empty block in the middle."""
    n, dpx, tiles = _three_tiles()
    _set(eng, [1.6, 3.2])
    masks = []
    for t in tiles:
        r, c = np.nonzero(np.triu(t, 4))
        masks.append((r, c, t[r, c]))
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0))
    order = [masks[0], empty, masks[1], masks[2]]
    eng.configure(n, dpx, 5)
    for b, m in enumerate(order):
        eng.upload_coo(b + 1, *m)
    eng.run()
    ref = eng.records_batch()
    eng.configure(n, dpx, 5)
    eng.upload_coo_batch(1, *blockrun.concat_coo(order))
    eng.run()
    got = eng.records_batch()
    assert ref[0]["n_found"] == 0 and ref[2]["n_found"] == 0 and ref[1]["n_found"] > 100
    for a, b in zip(ref, got):
        _equal_records(a, b)
    # and from COO that already lives on the device (mb200_upload_coo_dev: torch tensors as device buffers)
    import torch
    eng.configure(n, dpx, 5)
    for b, m in enumerate(order):
        dev = [torch.as_tensor(np.ascontiguousarray(a, dt), device="cuda:%d" % eng.device)
               for a, dt in zip(m, (np.int32, np.int32, np.float64))]
        torch.cuda.synchronize()
        eng.upload_coo_dev(b + 1, *dev)
    eng.run()
    for a, b in zip(ref, eng.records_batch()):
        _equal_records(a, b)


def _host_candidates(n, dpx, mask, rec, pt, st):
    """What mb200_select_candidates must deliver, derived on the host from the full record list with the product's
    (reference-pinned) sparse post-processing pieces."""
    pp = mm.postprocess
    mr, mc, mv = mask
    index = pp.MaskIndex(mr, mc, n)
    q = fdr_bh(rec["p"])
    sel = q < pt
    x, y, sg = rec["rows"][sel].astype(np.int64), rec["cols"][sel].astype(np.int64), rec["sigma"][sel]
    keep = pp.sparsity_filter(index, x, y, sg, st)
    pos = index.lookup(rec["rows"], rec["cols"])
    o_mask, s_mask = np.full(index.keys.size, 2.0), np.ones(index.keys.size)
    o_mask[pos], s_mask[pos] = q, rec["sigma"]

    def dense(vals, r, c):
        ok = (r >= 0) & (r < n) & (c >= 0) & (c < n)
        p_ = index.lookup(np.where(ok, r, 0), np.where(ok, c, 0))
        return np.where(ok & (p_ >= 0), vals[np.maximum(p_, 0)], 1.0)
    dr, dc = np.repeat([-1, 0, 1], 3), np.tile([-1, 0, 1], 3)
    o9 = np.stack([dense(o_mask, x + a, y + b) for a, b in zip(dr, dc)], axis=1).reshape(-1, 9)
    so9 = np.stack([dense(s_mask, x + a, y + b) for a, b in zip(dr, dc)], axis=1).reshape(-1, 9)
    d = y - x
    cval = np.where((d <= 4) | (d >= dpx + 1), 2.0, mv[np.maximum(index.lookup(x, y), 0)])
    return dict(rows=x, cols=y, q=q[sel], sigma=sg, keep=keep, cval=cval, o9=o9, so9=so9)


@pytest.mark.parametrize("octs,pt,st", [([1.6, 3.2], 0.3, 0.6), ([1.6, 3.2, 6.4, 12.8], 0.8, 0.3)])
def test_device_candidates_equal_host_selection(eng, octs, pt, st):
    """BH, o < pt, sparsity windows (incl. the negative-slice and clipped-window quirks near the tile edges) and the 3 x 3
    neighbourhoods of o / so: device vs the host's sparse post-processing on the same records, field for field."""
    n, dpx = 400, 150
    tiles = [gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=170 + b, blob_seed=180 + b, nblobs=25, missing=0.05 + 0.22 * b), n)
             for b in range(3)]
    _set(eng, octs)
    masks = []
    for t in tiles:
        r, c = np.nonzero(np.triu(t, 4))
        masks.append((r, c, t[r, c]))
    eng.configure(n, dpx, 3)
    eng.upload_coo_batch(0, *blockrun.concat_coo(masks))
    eng.run()
    recs = eng.records_batch()
    eng.select_candidates(pt, st)
    cands = eng.candidates_batch()
    seen_keep = seen_drop = seen_pass = seen_fail = False
    for b in range(3):
        assert np.array_equal(eng.q_values(b), fdr_bh(recs[b]["p"]))
        ref = _host_candidates(n, dpx, masks[b], recs[b], pt, st)
        got = cands[b]
        assert len(ref["rows"]) > 20
        seen_keep |= bool(ref["keep"].any())
        seen_drop |= bool((~ref["keep"]).any())
        for k in ("rows", "cols", "q", "sigma", "keep", "cval", "o9", "so9"):
            assert np.array_equal(got[k], ref[k]), (b, k)
        # enrichment filter (mustache.py:816-828) decided on the device: np.mean of the diagonal's non-zero entries bit for bit
        kx, ky = ref["rows"][ref["keep"]], ref["cols"][ref["keep"]]
        means = mm.postprocess.diagonal_means(*masks[b], dpx, ky - kx)
        with np.errstate(invalid="ignore"):
            passing = ref["cval"][ref["keep"]] > 2 * np.array([means[int(k)] for k in ky - kx])
        assert np.array_equal(got["enriched"][got["keep"]], passing), b
        assert not got["enriched"][~got["keep"]].any()
        seen_pass |= bool(passing.any())
        seen_fail |= bool((~passing).any())
    assert seen_keep and seen_drop                       # the sparsity filter both keeps and rejects on these tiles
    assert seen_pass and seen_fail                       # and so does the enrichment filter
    # a scratch pool of one diagonal forces one round per needed diagonal: same flags
    os.environ["MB200_ENRICH_POOL_KB"] = "1"
    try:
        eng.select_candidates(pt, st)
        again = eng.candidates_batch()
    finally:
        del os.environ["MB200_ENRICH_POOL_KB"]
    assert all(np.array_equal(again[b]["enriched"], cands[b]["enriched"]) for b in range(3))
    # candidate capacity overflow: the fetch re-runs the selection with room for every record
    eng.select_candidates(pt, st, candidate_fraction=1e-9)
    small = eng.candidates_batch()
    assert all(np.array_equal(small[b]["rows"], cands[b]["rows"]) for b in range(3))


@pytest.mark.parametrize("n,dpx,octs", [(512, 200, (1.6, 3.2)), (700, 900, (1.6, 3.2)), (2000, 400, (1.6, 3.2)),
                                        (1300, 500, (1.6, 3.2, 6.4, 12.8)), (640, 700, (1.6, 3.2, 6.4, 12.8))])
def test_fused_equals_three_kernel_path(eng, n, dpx, octs):
    """khs_kernel (axis-1 + DoG + scoring fused, DoG levels in shared memory) against kh_kernel + ks_kernel: records and
    coordinates, responses and scales bit for bit, p-values to 1e-12 (the two kernels sum |L| for the exponential fit in a
    different order), on band-limited tiles, a tile whose band is wider than the tile, and the CLI block shape."""
    tiles = [gen.band_to_dense(gen.dense_band_tile(n, min(dpx, n), seed=270 + b, blob_seed=280 + b, nblobs=20, missing=0.1 * b), n)
             for b in range(2)]
    _set(eng, list(octs))                                 # two octaves: 64-column tiles; four: 128-column tiles, one CTA per SM
    out = {}
    for mode in (1, 2, 0):                                # 1: khs_kernel, 2: kvh_kernel + ks_kernel, 0: kv + kh + ks
        eng.set_fusion(mode)
        eng.configure(n, dpx, 2)
        for b, t in enumerate(tiles):
            eng.upload_dense(b, t)
        eng.run()
        out[mode] = (eng.records_batch(), [eng.fits(b) for b in range(2)], eng.timing())
    assert out[1][2]["ks_ms"] < 0.02 < out[0][2]["ks_ms"]          # the paths really are different kernels
    if len(octs) == 2:
        assert out[2][2]["kh_ms"] < 0.02 < out[0][2]["kh_ms"]      # kvh_kernel is timed in the axis-0 slot
    for b in range(2):
        assert out[1][0][b]["n_found"] > 100
        _equal_records(out[1][0][b], out[0][0][b], keys=("rows", "cols", "v", "score_id", "sigma"))
        assert np.abs(out[1][0][b]["p"] - out[0][0][b]["p"]).max() <= 1e-12
        assert np.array_equal(out[1][1][b]["loc"], out[0][1][b]["loc"])
        assert np.abs(out[1][1][b]["scale"] / out[0][1][b]["scale"] - 1).max() <= 1e-13
        _equal_records(out[2][0][b], out[0][0][b])                   # same scoring kernel: everything bit for bit
        assert np.array_equal(out[2][1][b]["scale"], out[0][1][b]["scale"])


def test_overlapped_passes_equal_single_stream(eng):
    """mb200_set_overlap: half-batches pipelined on two streams (own halves of the scratch) give the records of the
    single-stream run bit for bit, for an odd number of blocks too."""
    n, dpx = 512, 200
    tiles = [gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=370 + b, blob_seed=380 + b, nblobs=12, missing=0.05 * b), n)
             for b in range(5)]
    _set(eng, [1.6, 3.2])
    out = {}
    for overlap in (False, True):
        eng.set_overlap(overlap)
        eng.configure(n, dpx, 5)
        for b, t in enumerate(tiles):
            eng.upload_dense(b, t)
        eng.run()
        out[overlap] = (eng.records_batch(), [eng.fits(b) for b in range(5)])
    eng.set_overlap(False)
    for b in range(5):
        assert out[True][0][b]["n_found"] > 100
        _equal_records(out[True][0][b], out[False][0][b])
        assert np.array_equal(out[True][1][b]["scale"], out[False][1][b]["scale"])
