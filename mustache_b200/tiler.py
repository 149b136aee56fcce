"""Block geometry and per-block inputs (host side).  Restates the tiling half of regulator()
(mustache.py:892-924) and the overlap de-duplication of process_block() (mustache.py:945-960)."""
import math

import numpy as np


def distance_in_px(distance_in_bp, res):
    return int(math.ceil(distance_in_bp // res))          # mustache.py:892


def block_geometry(n, dpx):
    """Overlapping CHUNK x CHUNK blocks: CHUNK = max(2*dpx, 2000), overlap = dpx, last block right-aligned
    (mustache.py:896-910).  Returns (chunk, starts, ends)."""
    chunk = max(2 * dpx, 2000)
    if n <= chunk:
        return chunk, [0], [n]
    starts, ends = [0], [chunk]
    while ends[-1] < n:
        starts.append(ends[-1] - dpx)
        ends.append(starts[-1] + chunk)
    ends[-1] = n
    starts[-1] = ends[-1] - chunk
    return chunk, starts, ends


def block_mask_size(i, starts, ends, overlap):
    """mustache.py:948-953."""
    if i == 0:
        return -1
    if i == len(starts) - 1:
        return ends[i - 1] - starts[i]
    return overlap


def keep_after_overlap(loop, start, mask_size):
    """mustache.py:958."""
    return loop[0] >= start + mask_size or loop[1] >= start + mask_size


def block_coo(x, y, v, start, end):
    """Entries of the block, block-local coordinates, original order (mustache.py:919-922)."""
    sel = (x >= start) & (x < end) & (y >= start) & (y < end)
    return x[sel] - start, y[sel] - start, v[sel]


class BlockSlicer:
    """block_coo for many blocks of one chromosome: one stable sort by row (none when the reader's output is already row
    sorted), then every block is a contiguous row range filtered by column -- O(block) instead of O(nnz) per block.
    Relative order of the entries is preserved, so duplicate coordinates still resolve last-write-wins."""

    def __init__(self, x, y, v):
        x, y, v = np.asarray(x), np.asarray(y), np.asarray(v)
        if x.size > 1 and not (x[1:] >= x[:-1]).all():
            order = np.argsort(x, kind="stable")
            x, y, v = x[order], y[order], v[order]
        self.x, self.y, self.v = x, y, v

    def block(self, start, end):
        lo, hi = np.searchsorted(self.x, [start, end], side="left")
        xs, ys, vs = self.x[lo:hi], self.y[lo:hi], self.v[lo:hi]
        sel = (ys >= start) & (ys < end)
        return xs[sel] - start, ys[sel] - start, vs[sel]


def block_mask_pixels(xc, yc, vc, chunk):
    """Mask pixels of the dense tile `cc[xc, yc] = vc` (mustache.py:923-924 + :699): last write wins for duplicate
    coordinates, value != 0, j - i >= 4.  Returned in row-major order."""
    key = xc.astype(np.int64) * chunk + yc.astype(np.int64)
    if key.size < 2 or (key[1:] > key[:-1]).all():            # already row-major and unique (the usual reader output)
        r, c, val = xc.astype(np.int64), yc.astype(np.int64), vc
        ok = (val != 0) & (c - r >= 4)
        return r[ok], c[ok], val[ok]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    last = np.ones(ks.size, dtype=bool)
    last[:-1] = ks[1:] != ks[:-1]
    idx = order[last]
    r, c, val = xc[idx].astype(np.int64), yc[idx].astype(np.int64), vc[idx]
    ok = (val != 0) & (c - r >= 4)
    return r[ok], c[ok], val[ok]
