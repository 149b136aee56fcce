"""GPU end-to-end parity: the product's mustache()/diff_mustache()/CLI against outputs of the unmodified reference."""
import os

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
G = synth.GOLDEN
FDR_TOL = 1e-6          # BASELINE.json north_star: p-values within 1e-6


def _read_tsv(path):
    lines = open(path).read().splitlines()
    assert lines[0] == "BIN1_CHR\tBIN1_START\tBIN1_END\tBIN2_CHROMOSOME\tBIN2_START\tBIN2_END\tFDR\tDETECTION_SCALE"
    rows = [l.split("\t") for l in lines[1:]]
    rows.sort(key=lambda r: (int(r[1]), int(r[4])))
    return rows


def _same_tsv(got_path, ref_path):
    got, ref = _read_tsv(got_path), _read_tsv(ref_path)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g[:6] == r[:6]                       # chromosome strings and bin coordinates: exact
        assert g[7] == r[7]                         # DETECTION_SCALE string: exact
        assert abs(float(g[6]) - float(r[6])) <= FDR_TOL
    return len(got)


@pytest.mark.parametrize("name", list(synth.SYNTH_TILES))
def test_mustache_function_matches_reference(name):
    from mustache_b200 import mustache as mm
    spec = synth.SYNTH_TILES[name]
    z = np.load(os.path.join(G, "synth_%s.npz" % name))
    c = synth.make_tile(**spec["gen"])
    loops = mm.mustache(c, "1", "1", 5000, [], 0, c.shape[0], -1, spec["dpx"], list(spec["octaves"]), spec["st"], spec["pt"])
    got, ref = np.array(loops, float).reshape(-1, 4), z["loops"]
    assert got.shape == ref.shape and len(ref) > 0
    assert np.array_equal(got[:, [0, 1, 3]], ref[:, [0, 1, 3]])
    assert np.abs(got[:, 2] - ref[:, 2]).max() <= FDR_TOL
    assert c[0, 0] == 2 and c[10, 10 + spec["dpx"] + 1 if 10 + spec["dpx"] + 1 < c.shape[0] else 0] == 2   # fills applied in place


def test_diff_engine_records_and_function():
    from mustache_b200 import diff_mustache as dm
    from mustache_b200.mustache import get_engine
    spec = synth.SYNTH_DIFF
    z = np.load(os.path.join(G, "diff_synth.npz"))
    c1, c2 = synth.make_pair(**spec["gen"])
    eng = get_engine()
    dm._set_octaves_diff(eng, spec["octaves"])
    eng.configure(c1.shape[0], spec["dpx"], 2)
    eng.upload_dense(0, c1)
    eng.upload_dense(1, c2)
    eng.run_differential()
    for b, pre in ((0, "m1_"), (1, "m2_")):
        r = eng.records(b, pair=True)
        assert np.array_equal(r["rows"], z[pre + "rows"]) and np.array_equal(r["cols"], z[pre + "cols"])
        assert np.array_equal(r["v"], z[pre + "v"]) and np.array_equal(r["sigma"], z[pre + "scale"])
        assert np.abs(r["p"] - z[pre + "p"]).max() <= 1e-12
        assert np.abs(r["pair"] - z[pre + "pair"]).max() <= 1e-9
    out = dm.diff_mustache(c1.copy(), c2.copy(), "1", "1", 5000, 0, c1.shape[0], -1, spec["dpx"], list(spec["octaves"]),
                           spec["st"], spec["pt"], spec["pt2"])
    for got, key in zip(out, ("loops1", "diff1", "loops2", "diff2")):
        ref = z[key]
        got = np.array(got, float).reshape(-1, 4)
        assert got.shape == ref.shape
        assert np.array_equal(got[:, [0, 1, 3]], ref[:, [0, 1, 3]]) and np.abs(got[:, 2] - ref[:, 2]).max() <= FDR_TOL


def test_cli_chr21_matches_reference_tsv(tmp_path):
    """README.md:49-51 command through the product CLI: 90 loops, coordinates and scale exact, FDR within 1e-6."""
    from mustache_b200 import mustache as mm
    raw, kr = synth.write_chr21_text(str(tmp_path))
    out = str(tmp_path / "chr21_out.tsv")
    mm.main(["-f", raw, "-b", kr, "-ch", "21", "-r", "5kb", "-pt", "0.1", "-st", "0.8", "-o", out])
    assert _same_tsv(out, os.path.join(G, "chr21_loops.tsv")) == 90


def test_cli_diff_chr21_matches_reference(tmp_path):
    """chr21 vs its binomial(0.6) thinning through the differential CLI ((163,18) loops, (141,6) diff)."""
    from mustache_b200 import diff_mustache as dm
    if not os.path.exists(os.path.join(G, "chr21_diff_loop1.tsv")):
        pytest.skip("differential CLI golden not generated")
    raw, kr = synth.write_chr21_text(str(tmp_path))
    thin = synth.write_chr21_thinned(str(tmp_path))
    out = str(tmp_path / "diff")
    dm.main(["-f1", raw, "-f2", thin, "-b1", kr, "-b2", kr, "-ch", "21", "-r", "5kb", "-pt", "0.05", "-pt2", "0.1",
             "-st", "0.8", "-o", out])
    for suf in ("loop1", "loop2", "diffloop1", "diffloop2"):
        _same_tsv(out + "." + suf, os.path.join(G, "chr21_diff_%s.tsv" % suf))
