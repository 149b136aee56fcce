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
