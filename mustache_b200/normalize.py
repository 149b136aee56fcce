"""Per-diagonal z-score normalisation of the sparse contact list (host side, numpy).

`normalize_sparse` restates mustache.py:622-686 in numpy, numerically identical to the reference (same np.convolve
calls on the same sequences, same np.mean/np.std element order): it is what the parity fixtures were produced with.
`normalize_sparse_device` is the same function on the GPU (SURVEY.md section 8(f) item 1, C ABI
mb200_normalize_sparse): np.mean / np.std bit for bit, the 2 Mb box sums to ~1e-13 (np.convolve's BLAS summation order
is CPU dependent, so no implementation can match it bit for bit on every host).  On bias-corrected maps the two agree
to ~1e-13 and give the same loops (chr21: 90 loops, FDR within 1e-6).  On raw integer counts (what diff_mustache.py
feeds for map 1, quirk #13) the reference's z-scores inside windows of identical counts are rounding noise divided by
rounding noise (val - mean ~ 1e-16, std ~ 1e-8), some of them NaN -> 0 -> off the mask: there the reference's own result
depends on its BLAS build, and only the numpy path on the same host reproduces it.  The CLI therefore uses the device
normaliser for bias-corrected input (`-b`) and the numpy one for raw counts; MUSTACHE_NORMALIZE=host|device overrides.
"""
import ctypes as C
import os
import math
import warnings

import numpy as np

LOCAL_WINDOW_BP = 2000000     # mustache.py:628, 631
MIN_LOCAL_COUNT = 30          # mustache.py:657-658


def _nan_to(value, fallback):
    return fallback if math.isnan(value) else value


def _by_diagonal(dist, ndiag):
    """Index lists of the contacts on diagonals 0..ndiag-1, each in the order a boolean mask `dist == d` selects them
    (ascending position), from ONE stable sort instead of ndiag passes over all contacts."""
    order = np.argsort(dist, kind="stable")
    bounds = np.searchsorted(dist[order], np.arange(ndiag + 1), side="left")
    return [order[bounds[d]:bounds[d + 1]] for d in range(ndiag)]


def normalize_sparse(x, y, v, resolution, distance_in_px):
    """In-place normalisation of `v`; returns the per-diagonal weights list the reference also returns (unused).

    Same arithmetic on the same arrays in the same order as mustache.py:622-686 (bit-identical output, pinned by
    tests/test_host_pipeline.py against a dump of the reference); only the selection of a diagonal's contacts differs:
    the reference builds `distances == d` for every d (O(nnz * dpx), 80 % of its 3 s on chr21), here one stable sort."""
    x = np.asarray(x)
    y = np.asarray(y)
    n = int(max(x.max(), y.max())) + 1                               # mustache.py:623
    weights = []
    dist = np.abs(y - x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        if (n - distance_in_px) * resolution > LOCAL_WINDOW_BP:
            box = np.ones(int(LOCAL_WINDOW_BP / resolution))
            for d, on_diag in enumerate(_by_diagonal(dist, 2 + distance_in_px)):
                rows = x[on_diag]
                line = np.zeros(n - d)
                line[rows] = v[on_diag] + 0.001                       # mustache.py:635
                if line.size == 0:
                    continue
                g_std = _nan_to(np.std(v[on_diag]), 1)
                g_mean = _nan_to(np.mean(v[on_diag]), 0)
                cnt = np.convolve(line != 0, box, mode="same")
                s1 = np.convolve(line, box, mode="same")
                s2 = np.convolve(line ** 2, box, mode="same")
                var = (s2 - s1 ** 2 / cnt) / (cnt - 1)                # mustache.py:650
                g_var = g_std ** 2
                np.nan_to_num(var, copy=False, neginf=g_var, posinf=g_var, nan=g_var)
                mu = s1 / cnt
                sparse_window = cnt < MIN_LOCAL_COUNT
                mu[sparse_window] = g_mean
                var[sparse_window] = g_var
                np.nan_to_num(mu, copy=False, neginf=g_mean, posinf=g_mean, nan=g_mean)
                sd = np.sqrt(var)
                line[rows] -= mu[rows]
                line[rows] /= sd[rows]
                np.nan_to_num(line, copy=False, nan=0, posinf=0, neginf=0)
                w = 1 + math.log(1 + g_mean, 30)                      # mustache.py:667
                line = line * w
                weights.append(w)
                v[on_diag] = line[rows]
        else:
            np.nan_to_num(v, copy=False, neginf=0, posinf=0, nan=0)
            for on_diag in _by_diagonal(dist, min(distance_in_px, n)):    # mustache.py:674-675 (not 2+dpx)
                g_std = _nan_to(np.std(v[on_diag]), 1)
                g_mean = _nan_to(np.mean(v[on_diag]), 0)
                v[on_diag] = (v[on_diag] - g_mean) / g_std
                np.nan_to_num(v, copy=False, nan=0, posinf=0, neginf=0)
    return weights


def normalize_sparse_device(eng, x, y, v, resolution, distance_in_px):
    """normalize_sparse(x, y, v, resolution, distance_in_px) on the engine's GPU; `v` is normalised in place."""
    from .engine import EngineError, _f64p, _i32p
    xs = np.ascontiguousarray(x, dtype=np.int32)
    ys = np.ascontiguousarray(y, dtype=np.int32)
    vv = v if (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous) else np.ascontiguousarray(v, dtype=np.float64)
    cap = int(distance_in_px) + 2
    w = np.zeros(cap)
    nw = C.c_int(0)
    st = eng.lib.mb200_normalize_sparse(eng.h, xs.ctypes.data_as(_i32p), ys.ctypes.data_as(_i32p), vv.ctypes.data_as(_f64p),
                                        len(vv), int(resolution), int(distance_in_px), w.ctypes.data_as(_f64p), cap,
                                        C.byref(nw))
    if st:
        raise EngineError(st, eng.lib.mb200_last_error(eng.h).decode())
    if vv is not v:
        v[:] = vv
    return list(w[:nw.value])


def normalize(x, y, v, resolution, distance_in_px, eng=None, biased=False):
    """What the CLI calls.  Bias-corrected maps (`-b`, the recommended way to run the reference) go through the device
    normaliser; raw counts stay on the numpy path: inside a window of identical counts the reference's z-score is rounding
    noise over rounding noise (some of it NaN -> 0 -> off the mask) and depends on the BLAS behind np.convolve, so only
    numpy on the same host reproduces it (diff_mustache.py feeds map 1 that way), and integer arrays
    must take the truncating numpy assignment (mustache.py:668).  MUSTACHE_NORMALIZE=host|device overrides."""
    mode = os.environ.get("MUSTACHE_NORMALIZE", "auto")
    integer_counts = np.asarray(v).dtype.kind in "iu"
    use_device = eng is not None and not integer_counts and (mode == "device" or (mode == "auto" and biased))
    if not use_device:
        return normalize_sparse(x, y, v, resolution, distance_in_px)
    return normalize_sparse_device(eng, x, y, v, resolution, distance_in_px)
