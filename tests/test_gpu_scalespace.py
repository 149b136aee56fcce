"""GPU parity tests proper: the CUDA scale-space engine, called through the C ABI, against the oracle and against
dumps of the unmodified reference.  Bit-exact for coordinates, detection scale, Gaussian/DoG values and vAll;
p-values within 1e-12 (the device sums |L| in a different order than numpy's pairwise mean)."""
import os

import numpy as np
import pytest

from mustache_b200 import ladder, tiler
from oracle import scalespace as osc
from tests import synth

pytestmark = pytest.mark.gpu
G = synth.GOLDEN
P_TOL = 1e-12


@pytest.fixture(scope="module")
def eng():
    from mustache_b200.engine import ScaleSpaceEngine
    e = ScaleSpaceEngine(0)
    yield e
    e.close()


def _band_mask(n, lo, hi):
    """Pixels the detector reads: the scored band (diagonals lo..hi) dilated by the 3x3 maximum footprint."""
    from scipy.ndimage import binary_dilation
    d = np.subtract.outer(np.arange(n), np.arange(n)) * -1
    return binary_dilation((d >= lo) & (d <= hi), structure=np.ones((3, 3), bool))


@pytest.mark.parametrize("name", ["n256_o2", "n320_o4", "n200_full"])
def test_gaussian_and_dog_levels_bit_exact(eng, name):
    spec = synth.SYNTH_TILES[name]
    c = synth.make_tile(**spec["gen"])
    n = c.shape[0]
    prog = ladder.build_program(list(spec["octaves"]))
    eng.set_program(prog)
    eng.configure(n, spec["dpx"], 1)
    eng.upload_dense(0, c)
    nz, filled = osc.mask_and_fill(c, spec["dpx"])
    dhi = min(spec["dpx"] + 1, n - 1)
    band = _band_mask(n, 4, dhi)
    prev = None
    for s, st in enumerate(prog.steps):
        ref = osc.gaussian_level(filled, st.taps)
        if s % 3 == 0 or s == len(prog.steps) - 1 or st.restart:
            g, l = eng.debug_level(0, s)
            assert np.array_equal(g[band], ref[band]), "Gaussian step %d (sigma %.4f)" % (s, st.sigma)
            if not st.restart:
                assert np.array_equal(l[band], (prev - ref)[band]), "DoG step %d" % s
        prev = ref


def _check_records(rec, z, prefix="", sigma_exact=True):
    assert rec["nz_count"] == int(z[prefix + "nz_count"])
    assert np.array_equal(rec["rows"], z[prefix + "rows"])
    assert np.array_equal(rec["cols"], z[prefix + "cols"])
    assert np.array_equal(rec["v"], z[prefix + "v"])
    assert np.array_equal(rec["sigma"], z[prefix + "scale"])
    assert np.abs(rec["p"] - z[prefix + "p"]).max() <= P_TOL


@pytest.mark.parametrize("name", list(synth.SYNTH_TILES))
@pytest.mark.parametrize("dedupe", [True, False])
def test_records_match_reference_dump_synthetic(eng, name, dedupe):
    spec = synth.SYNTH_TILES[name]
    z = np.load(os.path.join(G, "synth_%s.npz" % name))
    c = synth.make_tile(**spec["gen"])
    eng.set_octaves(spec["octaves"], dedupe=dedupe)
    rec = eng.scale_space_dense(c, spec["dpx"])
    _check_records(rec, z)


def test_upload_paths_agree(eng):
    """dense host, dense device (torch tensor as the buffer), COO and band uploads give identical records."""
    import torch
    from mustache_b200 import synth as gen
    spec = synth.SYNTH_TILES["n256_o2"]
    band = gen.dense_band_tile(**{k: (min(v, spec["gen"]["n"]) if k == "dpx" else v) for k, v in spec["gen"].items()})
    c = gen.band_to_dense(band, 256)
    eng.set_octaves(spec["octaves"])
    eng.configure(256, spec["dpx"], 4)
    eng.upload_dense(0, c)
    eng.upload_dense(1, torch.from_numpy(c).cuda())
    r, cc, v = gen.band_to_coo(band, 256)
    eng.upload_coo(2, r, cc, v)
    eng.upload_band(3, np.ascontiguousarray(band))
    eng.run()
    recs = [eng.records(b) for b in range(4)]
    z = np.load(os.path.join(G, "synth_n256_o2.npz"))
    for rec in recs:
        _check_records(rec, z)


@pytest.fixture(scope="module")
def chr21(tmp_path_factory):
    from mustache_b200 import normalize, readers
    d = tmp_path_factory.mktemp("chr21")
    raw, kr = synth.write_chr21_text(str(d))
    x, y, v = readers.read_text(raw, 2000000, kr, "21", 5000)
    dpx = tiler.distance_in_px(2000000, 5000)
    normalize.normalize_sparse(x, y, v, 5000, dpx)
    return x, y, v, dpx


def test_chr21_blocks_match_reference_dump(eng, chr21):
    """All six chr21 blocks in one batch (COO upload) against what the reference's mustache() held per block."""
    x, y, v, dpx = chr21
    z = np.load(os.path.join(G, "chr21_blocks.npz"))
    n = int(max(x.max(), y.max()) + 1)
    chunk, starts, ends = tiler.block_geometry(n, dpx)
    eng.set_octaves([1.6, 3.2])
    eng.configure(chunk, dpx, len(starts))
    for b, (s, e) in enumerate(zip(starts, ends)):
        xc, yc, vc = tiler.block_coo(x, y, v, s, e)
        mr, mc, mv = tiler.block_mask_pixels(xc, yc, vc, chunk)
        eng.upload_coo(b, mr, mc, mv)
    eng.run()
    for b in range(len(starts)):
        rec = eng.records(b)
        assert rec["nz_count"] == int(z["b%d_nz_count" % b])
        if b == 0:
            continue            # 1 mask pixel: the reference returns before the loop (mustache.py:701)
        _check_records(rec, z, "b%d_" % b)
    t = eng.timing()
    assert t["total_ms"] > 0 and eng.launches() >= 4


def test_fits_match_oracle(eng):
    spec = synth.SYNTH_TILES["n256_o2"]
    c = synth.make_tile(**spec["gen"])
    res = osc.scale_space(c, spec["dpx"], spec["octaves"])
    eng.set_octaves(spec["octaves"])
    eng.configure(256, spec["dpx"], 1)
    eng.upload_dense(0, c)
    eng.run()
    f = eng.fits(0)
    ref = {o * 12 + i: (loc, sc) for o, i, loc, sc in res["fits"]}
    for sid, loc, sc in zip(f["score_id"], f["loc"], f["scale"]):
        assert loc == ref[int(sid)][0]
        assert abs(sc - ref[int(sid)][1]) <= 1e-13 * abs(sc)


def test_error_codes(eng):
    from mustache_b200.engine import EngineError
    eng.set_octaves([1.6, 3.2])
    with pytest.raises(EngineError) as ei:
        eng.configure(20, 10, 1)              # tile smaller than the filter support
    assert ei.value.code == -2
    eng.configure(128, 50, 1)
    bad = np.zeros((128, 128))
    bad[10, 30] = np.nan
    bad[5:100, 20:120] += np.triu(np.ones((95, 100)), 6)
    eng.upload_dense(0, bad)
    eng.run()
    with pytest.raises(EngineError) as ei:
        eng.records(0)
    assert ei.value.code == -4



def test_records_device_alias_matches_host_fetch(eng):
    """records_device() hands NCCL the engine's own buffers (no host round trip): same content as the host fetch."""
    spec = synth.SYNTH_TILES["n256_o2"]
    c = synth.make_tile(**spec["gen"])
    eng.set_octaves(spec["octaves"])
    eng.configure(256, spec["dpx"], 1)
    eng.upload_dense(0, c)
    eng.run()
    host = eng.records(0, sort=False)
    dev = eng.records_device(0)
    assert dev["n_found"] == host["n_found"] and dev["nz_count"] == host["nz_count"]
    assert np.array_equal(dev["rows"].cpu().numpy(), host["rows"]) and np.array_equal(dev["cols"].cpu().numpy(), host["cols"])
    assert np.array_equal(dev["v"].cpu().numpy(), host["v"]) and np.array_equal(dev["p"].cpu().numpy(), host["p"])
    ids = eng.fits(0)["score_id"][dev["scored_index"].cpu().numpy()]
    assert np.array_equal(ids, host["score_id"])


def test_pipelined_uploads_do_not_mix_batches(eng):
    """Uploads of batch k+1 are issued while batch k may still be running (double-buffered tiles): results of both
    batches must be those of their own tiles."""
    a = synth.make_tile(**synth.SYNTH_TILES["n256_o2"]["gen"])
    b = synth.make_tile(n=256, dpx=100, seed=77, blob_seed=78, nblobs=9, missing=0.2)
    eng.set_octaves([1.6, 3.2])
    eng.configure(256, 100, 1)
    ref = {}
    for name, t in (("a", a), ("b", b)):
        eng.upload_dense(0, t)
        eng.run()
        ref[name] = eng.records(0)
    eng.upload_dense(0, a)
    eng.run()
    eng.upload_dense(0, b)          # goes to the other slot while the run of `a` is in flight
    ra = eng.records(0)
    eng.run()
    eng.upload_dense(0, a)
    rb = eng.records(0)
    eng.run()
    ra2 = eng.records(0)
    for got, want in ((ra, ref["a"]), (rb, ref["b"]), (ra2, ref["a"])):
        assert np.array_equal(got["rows"], want["rows"]) and np.array_equal(got["cols"], want["cols"])
        assert np.array_equal(got["v"], want["v"]) and np.array_equal(got["p"], want["p"])


@pytest.mark.parametrize("octaves", [(0.5,), (0.7, 1.4, 2.8), (2.3, 4.6), (1.6, 3.2, 6.4)])
def test_unusual_ladders_match_oracle(eng, octaves):
    """-sz / -oc other than the defaults: radii 1-2 (the axis-0 windows reach above the tile), one octave (a single
    chain), three octaves, a sigma0 whose octave tails are not bit-identical to the next octave's heads.  Records against
    the oracle (numpy restatement of mustache.py:699-772), Gaussians of every step bit-exact."""
    c = synth.make_tile(n=288, dpx=120, seed=71, blob_seed=72, nblobs=10, missing=0.1)
    prog = ladder.build_program(list(octaves))
    eng.set_program(prog)
    eng.configure(288, 120, 1)
    eng.upload_dense(0, c)
    nz, filled = osc.mask_and_fill(c, 120)
    band = _band_mask(288, 4, 121)
    for s, st in enumerate(prog.steps):
        g, _ = eng.debug_level(0, s)
        assert np.array_equal(g[band], osc.gaussian_level(filled, st.taps)[band]), "step %d radius %d" % (s, st.radius)
    eng.upload_dense(0, c)
    eng.run()
    rec = eng.records(0)
    ref = osc.scale_space(c, 120, list(octaves))
    found = ref["p"] != 2
    assert rec["nz_count"] == ref["nz_count"] and found.sum() > 0
    assert np.array_equal(rec["rows"], ref["rows"][found]) and np.array_equal(rec["cols"], ref["cols"][found])
    assert np.array_equal(rec["v"], ref["v"][found]) and np.array_equal(rec["sigma"], ref["scale"][found])
    assert np.abs(rec["p"] - ref["p"][found]).max() <= P_TOL


def test_two_engines_alternating_give_the_single_engine_results(eng):
    """mb200_run_after: two engine handles on one GPU used alternately for a stream of batches (the run of batch k+1 behind
    the run of batch k, post-processing and fetch of batch k next to it) return exactly what one engine returns batch by
    batch -- records, device-selected candidates and enrichment flags."""
    from mustache_b200 import synth as gen
    from mustache_b200.engine import ScaleSpaceEngine, EngineError
    n, dpx, octs = 512, 200, [1.6, 3.2]
    batches = [[gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=70 + 3 * k + b, blob_seed=170 + 3 * k + b, nblobs=12,
                                                     missing=0.3), n) for b in range(2)] for k in range(5)]

    def keyed(r):
        o = np.lexsort((r["cols"], r["rows"]))
        return {k: np.asarray(v)[o] for k, v in r.items() if isinstance(v, np.ndarray)}

    def same(a, b):
        a, b = keyed(a), keyed(b)
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k]), k

    eng.set_octaves(octs)
    ref = []
    for tiles in batches:
        eng.configure(n, dpx, 2)
        for b, t in enumerate(tiles):
            eng.upload_dense(b, t)
        eng.run()
        eng.select_candidates(0.2, 0.88)
        ref.append((eng.candidates_batch(), [eng.records(b) for b in range(2)]))
    assert sum(len(c["rows"]) for cands, _ in ref for c in cands) > 0

    other = ScaleSpaceEngine(eng.device)
    try:
        engines = [eng, other]
        for e in engines:
            e.set_octaves(octs)
            e.configure(n, dpx, 2)
        for k in range(2):                                   # first tiles of each engine
            for b, t in enumerate(batches[k]):
                engines[k].upload_dense(b, t)
        got = []
        for i in range(len(batches) + 1):
            e, o = engines[i % 2], engines[(i + 1) % 2]
            if i < len(batches):
                if i > 0:
                    e.run_after(o)
                e.run()
                if i + 2 < len(batches):                     # tiles of the batch this engine runs next
                    for b, t in enumerate(batches[i + 2]):
                        e.upload_dense(b, t)
            if i > 0:
                o.select_candidates(0.2, 0.88)
                got.append((o.candidates_batch(), [o.records(b) for b in range(2)]))
        assert len(got) == len(ref)
        for (gc, gr), (rc, rr) in zip(got, ref):
            for b in range(2):
                same(gc[b], rc[b])
                same(gr[b], rr[b])
                assert gc[b]["n_found"] == rc[b]["n_found"] and gc[b]["nz_count"] == rc[b]["nz_count"]
        with pytest.raises(EngineError):
            eng.run_after(eng)
    finally:
        other.close() if hasattr(other, "close") else None
