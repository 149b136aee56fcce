"""Dense CPU restatement of the reference's block post-processing.  TEST INFRASTRUCTURE ONLY (see oracle/scalespace.py).

Follows mustache.py:774-850 on dense N x N arrays exactly as the reference does (dense `o`, `so`, argsort,
raw-slice sparsity windows, vectorised diagonal means, float32 label image + scipy.ndimage.label), so that the
product's sparse implementation (mustache_b200/postprocess.py) can be checked against it and against the
reference's own outputs (tests/golden/*_loops).  statsmodels' fdr_bh is restated from fdrcorrection(method='indep').
"""
import math

import numpy as np
from scipy import ndimage


def bh(p):
    p = np.asarray(p, float)
    m = p.size
    srt = np.argsort(p)
    adj = p[srt] * m / np.arange(1, m + 1)
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    adj = np.minimum(adj, 1.0)
    out = np.empty(m)
    out[srt] = adj
    return out


def bh_statsmodels_form(p):
    """Same operation order as statsmodels (p / (rank/m)); used to pin bit-level agreement of the product's fdr_bh."""
    p = np.asarray(p, float)
    m = p.size
    srt = np.argsort(p)
    adj = np.take(p, srt) / (np.arange(1, m + 1) / float(m))
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    adj[adj > 1] = 1
    out = np.empty(m)
    out[srt] = adj
    return out


def loops_dense(filled, nz, p_all, scales, start, dpx, st, pt, intra=True):
    """filled: tile after the 2-fills; nz: mask; p_all/scales: per-mask-pixel state after the scale-space loop."""
    if p_all.size < 10000:                                         # mustache.py:775
        return []
    p_all = p_all.copy()
    found = p_all != 2
    p_all[found] = bh_statsmodels_form(p_all[found])               # mustache.py:778-779
    o = np.ones_like(filled)
    o[nz] = p_all
    so = np.ones_like(filled)
    so[nz] = scales
    k = int((o < pt).sum())
    flat = np.argsort(o, axis=None)[:k]
    xs, ys = np.unravel_index(flat, o.shape)
    ok = xs != 0                                                   # mustache.py:800
    for t in range(k):
        s = math.ceil(so[xs[t], ys[t]])
        d1 = nz[xs[t] - s:xs[t] + s + 1, ys[t] - s:ys[t] + s + 1].sum() / (2 * s + 1) ** 2
        s *= 2
        d2 = nz[xs[t] - s:xs[t] + s + 1, ys[t] - s:ys[t] + s + 1].sum() / (2 * s + 1) ** 2
        if d1 < st or d2 < 0.6:
            ok[t] = False
    xs, ys = xs[ok], ys[ok]
    if xs.size == 0:
        return []
    if intra:
        with np.errstate(all="ignore"):
            means = np.array([_diag_nonzero_mean(filled, int(dd)) for dd in ys - xs])
            good = filled[xs, ys] > 2 * means
        if good.sum() == 0:
            return []
        xs, ys = xs[good], ys[good]
    side = int(ys.max()) + 2
    img = np.zeros((side, side), dtype=np.float32)
    img[xs, ys] = o[xs, ys] + 1
    for da in (-1, 0, 1):
        for db in (-1, 0, 1):
            if da or db:
                img[xs + da, ys + db] = 2
    lab, ncomp = ndimage.label(img, structure=np.ones((3, 3)))
    out = []
    for c in range(1, ncomp + 1):
        pts = np.argwhere(lab == c)
        w = int(np.argmin(o[pts[:, 0], pts[:, 1]]))
        a, b = int(pts[w, 0]), int(pts[w, 1])
        out.append([a + start, b + start, float(o[a, b]), float(so[a, b])])
    return out


def _diag_nonzero_mean(m, k):
    d = np.diagonal(m, k)
    d = d[d != 0]
    return np.mean(d)
