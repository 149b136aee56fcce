"""Block post-processing on sparse records (host side; consumes the scale-space engine's output).

Restates mustache.py:774-850 (and the per-map half of diff_mustache.py:428-561) without ever materialising the
dense N x N `o` / `so` / label matrices, but reproducing their observable behaviour exactly, including the quirks
listed in SURVEY.md App. D (#2 mask-size guard, #3 `nonsparse = x != 0`, #4 raw numpy slice windows, #5 cluster
representative chosen over ALL component pixels).

Inputs are block-local:
  mask_rows, mask_cols   every mask pixel (c != 0 and j-i >= 4, mustache.py:699) in row-major order
  mask_vals              the normalised contact value at those pixels (before the 2-fill)
  rec_*                  records of the pixels the scale-space loop updated (p != 2): row, col, p_raw, sigma
"""
import math

import numpy as np

from .fdr import fdr_bh

MIN_MASK_FOR_BH = 10000   # mustache.py:775: `len(pFound)` is the MASK size


def _slice_bounds(lo, hi_plus1, n):
    """Python/numpy semantics of a[lo:hi_plus1] on an axis of length n (vectorised): returns (start, stop)."""
    lo = np.where(lo < 0, np.maximum(lo + n, 0), np.minimum(lo, n))
    hi = np.where(hi_plus1 < 0, np.maximum(hi_plus1 + n, 0), np.minimum(hi_plus1, n))
    return lo, hi


class MaskIndex:
    """Row-major sorted keys of the mask pixels; rectangle counts by binary search."""

    def __init__(self, rows, cols, n):
        self.n = int(n)
        self.keys = rows.astype(np.int64) * self.n + cols.astype(np.int64)
        if self.keys.size > 1 and not (np.diff(self.keys) > 0).all():
            order = np.argsort(self.keys, kind="stable")
            self.keys = self.keys[order]
            self.order = order
        else:
            self.order = None

    def window_counts(self, x, y, half):
        """sum(nz[x-h:x+h+1, y-h:y+h+1]) per candidate, with numpy's slice semantics (mustache.py:803-807)."""
        n = self.n
        r0, r1 = _slice_bounds(x - half, x + half + 1, n)
        c0, c1 = _slice_bounds(y - half, y + half + 1, n)
        out = np.zeros(len(x), dtype=np.int64)
        span = int((r1 - r0).max()) if len(x) else 0
        for k in range(max(span, 0)):
            r = r0 + k
            live = (r < r1) & (c0 < c1)
            if not live.any():
                continue
            base = r[live] * n
            lo = np.searchsorted(self.keys, base + c0[live], side="left")
            hi = np.searchsorted(self.keys, base + c1[live], side="left")
            out[live] += hi - lo
        return out

    def lookup(self, rows, cols):
        """Index of (row, col) in the sorted mask, or -1."""
        k = np.asarray(rows, np.int64) * self.n + np.asarray(cols, np.int64)
        pos = np.searchsorted(self.keys, k)
        pos = np.minimum(pos, self.keys.size - 1)
        hit = self.keys[pos] == k
        return np.where(hit, pos, -1)


def sparsity_filter(index, x, y, scales, st):
    """mustache.py:800-811.  Returns the boolean `nonsparse` vector."""
    keep = x != 0
    if len(x) == 0:
        return keep
    s = np.ceil(scales).astype(np.int64)
    c1 = index.window_counts(x, y, s) / ((2 * s + 1) ** 2)
    s2 = 2 * s
    c2 = index.window_counts(x, y, s2) / ((2 * s2 + 1) ** 2)
    keep &= ~((c1 < st) | (c2 < 0.6))
    return keep


def diagonal_means(mask_rows, mask_cols, mask_vals, dpx, wanted, intra=True):
    """mean of the non-zero entries of the k-th diagonal of the 2-FILLED tile (mustache.py:816-823).

    Diagonals k <= 4 and (intra) k >= dpx+1 are constant 2 after the fills (mustache.py:703-706); for the others the
    non-zero entries are exactly the mask pixels on that diagonal, gathered in row order so np.mean's pairwise
    summation sees the same sequence as the reference's `vals[vals != 0]`.
    """
    d = mask_cols.astype(np.int64) - mask_rows.astype(np.int64)
    out = {}
    uniq = np.unique(wanted)
    by_diag = None
    if len(uniq) > 16:                          # many diagonals: one stable sort instead of a pass over the mask per diagonal
        order = np.argsort(d, kind="stable")
        ds = d[order]
        by_diag = (order, np.searchsorted(ds, uniq, side="left"), np.searchsorted(ds, uniq, side="right"))
    for pos, k in enumerate(uniq):
        k = int(k)
        if k <= 4 or (intra and k >= dpx + 1):
            out[k] = 2.0
            continue
        vals = mask_vals[d == k] if by_diag is None else mask_vals[by_diag[0][by_diag[1][pos]:by_diag[2][pos]]]
        vals = vals[vals != 0]
        with np.errstate(all="ignore"):
            out[k] = float(np.mean(vals)) if vals.size else float("nan")
    return out


def cluster_representatives(x, y, o_of):
    """mustache.py:830-848 without the dense label matrix.

    Foreground = union of the 3x3 neighbourhoods of the candidates; 8-connected components; per component the
    pixel with the smallest `o` (first in row-major order on ties) over ALL its pixels.  Components are emitted in
    scipy.ndimage.label order (raster order of their first pixel).  `o_of(rows, cols)` returns the dense-`o` value.
    """
    cand = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(x, y))}
    parent = list(range(len(x)))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for (a, b), i in cand.items():
        for da in range(-3, 4):
            for db in range(-3, 4):
                j = cand.get((a + da, b + db))
                if j is not None and j != i:
                    ri, rj = find(i), find(j)
                    if ri != rj:
                        parent[rj] = ri
    comps = {}
    for (a, b), i in cand.items():
        comps.setdefault(find(i), set()).update((a + da, b + db) for da in (-1, 0, 1) for db in (-1, 0, 1))
    out = []
    for pix in comps.values():
        pts = np.array(sorted(pix), dtype=np.int64)          # row-major order == np.argwhere order
        vals = o_of(pts[:, 0], pts[:, 1])
        k = int(np.argmin(vals))
        out.append((tuple(pts[0]), int(pts[k, 0]), int(pts[k, 1])))
    out.sort(key=lambda t: t[0])
    return [(a, b) for _, a, b in out]


def call_loops(n, dpx, start, mask_rows, mask_cols, mask_vals, rec_rows, rec_cols, rec_p, rec_sigma, st, pt,
               intra=True, candidate_order="sorted", partial=False):
    """Everything after the scale-space loop for one block (mustache.py:774-850).

    Returns (loops, aux): loops = [[x+start, y+start, fdr, sigma], ...]; aux carries q (per record) and the
    lookup helpers the differential selection needs; aux["empty_after_filters"] is set when the sparsity or the
    enrichment filter left nothing (diff_mustache.py:507-508, 519-520, 526-527 abandon the whole block pair then).
    candidate_order: "sorted" = argsort(o) as mustache.py:792, "rowmajor" = np.where(o < pt) as diff_mustache.py:458;
    only the candidate SET influences the result.
    """
    aux = {}
    if len(mask_rows) < MIN_MASK_FOR_BH:
        return [], aux
    index = MaskIndex(mask_rows, mask_cols, n)
    if index.order is not None:
        mask_rows, mask_cols, mask_vals = mask_rows[index.order], mask_cols[index.order], mask_vals[index.order]
    q = fdr_bh(rec_p)
    aux.update(q=q, index=index)
    rec_pos = index.lookup(rec_rows, rec_cols)
    o_mask = np.full(index.keys.size, 2.0)       # on-mask, never updated: pAll stays 2 (mustache.py:708)
    s_mask = np.ones(index.keys.size)            # Scales initialised to 1 (mustache.py:709)
    o_mask[rec_pos] = q
    s_mask[rec_pos] = rec_sigma
    aux.update(o_mask=o_mask, s_mask=s_mask)

    def o_of(r, c):                               # dense `o`: 1 off-mask (mustache.py:789-790)
        pos = index.lookup(r, c)
        return np.where(pos >= 0, o_mask[np.maximum(pos, 0)], 1.0)

    def so_of(r, c):
        pos = index.lookup(r, c)
        return np.where(pos >= 0, s_mask[np.maximum(pos, 0)], 1.0)
    aux.update(o_of=o_of, so_of=so_of)

    sel = q < pt                                  # `o < pt`: off-mask (1) and untouched (2) never pass for pt <= 1
    if pt > 1:
        raise ValueError("pt > 1 would select off-mask pixels in the reference; not supported")
    x = rec_rows[sel].astype(np.int64)
    y = rec_cols[sel].astype(np.int64)
    sc = np.asarray(rec_sigma)[sel]
    if candidate_order == "sorted":               # argsort(o.ravel()) order (mustache.py:792); only the SET matters
        order = np.argsort(q[sel], kind="stable")
        x, y, sc = x[order], y[order], sc[order]
    keep = sparsity_filter(index, x, y, sc, st)
    x, y = x[keep], y[keep]
    if len(x) == 0:
        aux["empty_after_filters"] = True
        return [], aux
    if intra:
        d = y - x
        means = diagonal_means(mask_rows, mask_cols, mask_vals, dpx, d, intra)
        mvec = np.array([means[int(k)] for k in d])
        pos = index.lookup(x, y)
        cxy = np.where((d <= 4) | (d >= dpx + 1), 2.0, mask_vals[pos])
        with np.errstate(invalid="ignore"):
            passing = cxy > 2 * mvec
        if passing.sum() == 0:
            aux["empty_after_filters"] = True
            return [], aux
        x, y = x[passing], y[passing]
    reps = cluster_representatives(x, y, o_of)
    loops = []
    for a, b in reps:
        loops.append([a + start, b + start, float(o_of([a], [b])[0]), float(so_of([a], [b])[0])])
    return loops, aux


def call_loops_from_candidates(n, dpx, start, mask_rows, mask_cols, mask_vals, cand, intra=True, extra=None):
    """The rest of mustache.py:774-850 for one block when the device already did BH, the `o < pt` cut and the sparsity
    filter (mb200_select_candidates): `cand` is one entry of ScaleSpaceEngine.candidates_batch().  Enrichment filter
    (:816-828) from the block's mask pixels, clustering (:830-848) from the candidates' 3 x 3 neighbourhoods of `o` / `so`.
    Same loops as call_loops() on the full record list.  extra: names of further [m, 9] neighbourhood arrays of `cand`
    (pair9, vself9, vother9 of a differential run); the function then returns (loops, [{name: value at the loop}], emptied)
    where `emptied` tells that a filter left nothing (diff_mustache.py:507-527 abandons the whole block pair then)."""
    def done(loops, vals=(), emptied=False):
        return loops if extra is None else (loops, list(vals), emptied)
    if len(mask_rows) < MIN_MASK_FOR_BH:                     # mustache.py:775: `len(pFound)` is the MASK size
        return done([])
    keep = cand["keep"]
    x, y = cand["rows"][keep].astype(np.int64), cand["cols"][keep].astype(np.int64)
    if len(x) == 0:
        return done([], emptied=True)
    o9, so9 = cand["o9"][keep], cand["so9"][keep]
    more = {k: cand[k][keep] for k in (extra or ())}
    if intra:
        if "enriched" in cand:                                # decided on the device (mb200_enrich_candidates)
            passing = cand["enriched"][keep]
        else:
            d = y - x
            means = diagonal_means(np.asarray(mask_rows), np.asarray(mask_cols), np.asarray(mask_vals), dpx, d, intra)
            mvec = np.array([means[int(k)] for k in d])
            with np.errstate(invalid="ignore"):
                passing = cand["cval"][keep] > 2 * mvec
        if passing.sum() == 0:
            return done([], emptied=True)
        x, y, o9, so9 = x[passing], y[passing], o9[passing], so9[passing]
        more = {k: v[passing] for k, v in more.items()}
    o_at, so_at, more_at = {}, {}, {k: {} for k in more}
    for i, (a, b, ov, sv) in enumerate(zip(x.tolist(), y.tolist(), o9, so9)):
        for k in range(9):
            key = (a + k // 3 - 1, b + k % 3 - 1)
            o_at[key] = ov[k]
            so_at[key] = sv[k]
            for name, arr in more.items():
                more_at[name][key] = arr[i, k]

    def o_of(r, c):
        return np.array([o_at[(int(a), int(b))] for a, b in zip(r, c)])
    reps = cluster_representatives(x, y, o_of)
    loops = [[a + start, b + start, float(o_at[(a, b)]), float(so_at[(a, b)])] for a, b in reps]
    return done(loops, [{name: float(more_at[name][(a, b)]) for name in more} for a, b in reps])
