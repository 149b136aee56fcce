"""Differential selection (diff_mustache.py:428-569) restated on sparse records, against the reference's own dump."""
import os

import numpy as np

from mustache_b200 import diff_mustache as dm
from tests import synth

G = synth.GOLDEN


def _masks(c):
    n = c.shape[0]
    d = np.subtract.outer(np.arange(n), np.arange(n)) * -1
    r, cc = np.nonzero((c != 0) & (d >= 4))
    return r, cc, c[r, cc]


def test_select_differential_matches_reference():
    spec = synth.SYNTH_DIFF
    z = np.load(os.path.join(G, "diff_synth.npz"))
    c1, c2 = synth.make_pair(**spec["gen"])
    recs = []
    for pre in ("m1_", "m2_"):
        recs.append(dict(rows=z[pre + "rows"], cols=z[pre + "cols"], v=z[pre + "v"], p=z[pre + "p"], sigma=z[pre + "scale"],
                         pair=z[pre + "pair"], nz_count=int(z[pre + "nz_count"])))
    out = dm.select_differential(c1.shape[0], spec["dpx"], 0, [_masks(c1), _masks(c2)], recs, spec["st"], spec["pt"], spec["pt2"])
    for got, key in zip(out, ("loops1", "diff1", "loops2", "diff2")):
        ref = z[key]
        assert len(ref) > 0
        assert np.array_equal(np.array(got, float).reshape(-1, 4), ref), key


def test_select_differential_small_mask():
    r = np.arange(60)
    m = (r, r + 10, np.ones(60))
    rec = dict(rows=r[:3], cols=r[:3] + 10, v=np.ones(3), p=np.full(3, 1e-8), sigma=np.full(3, 2.1), pair=np.zeros(3), nz_count=60)
    assert dm.select_differential(2000, 400, 0, [m, m], [rec, rec], 0.0, 0.5, 0.5) == ([], [], [], [])


def test_candidate_based_differential_selection():
    """The split the product uses for diff_mustache (device: BH, o < pt, sparsity filter and the o / so / pair / v
    neighbourhoods per map; host: enrichment, clustering, pair < pt2 and v_self > v_other from those candidates) returns the
    reference's four loop lists.  The device half is stood in for by its dense restatement on the oracle."""
    from mustache_b200 import postprocess
    from tests.test_sharded_blocks import OracleEngine
    spec = synth.SYNTH_DIFF
    z = np.load(os.path.join(G, "diff_synth.npz"))
    c1, c2 = synth.make_pair(**spec["gen"])
    masks = [_masks(c1), _masks(c2)]
    eng = OracleEngine()
    eng.set_octaves(list(spec["octaves"]))
    eng.configure(c1.shape[0], spec["dpx"], 2)
    for b, m in enumerate(masks):
        eng.upload_coo(b, *m)
    eng.run_differential()
    eng.select_candidates(spec["pt"], spec["st"])
    cands = eng.candidates_batch(pair=True)
    out = dm.select_differential_from_candidates(c1.shape[0], spec["dpx"], 0, masks, cands, spec["pt2"])
    for got, key in zip(out, ("loops1", "diff1", "loops2", "diff2")):
        ref = z[key]
        got = np.array(got, float).reshape(-1, 4)
        assert got.shape == ref.shape and len(ref) > 0
        assert np.array_equal(got[:, [0, 1, 3]], ref[:, [0, 1, 3]]) and np.abs(got[:, 2] - ref[:, 2]).max() <= 1e-9
