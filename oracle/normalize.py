"""Literal CPU restatement of the reference's normalize_sparse (mustache.py:622-686): one boolean mask per diagonal,
exactly as the reference selects contacts.  TEST INFRASTRUCTURE ONLY: pins mustache_b200/normalize.py (which selects the
same contacts from one stable sort) bit for bit; itself pinned against a dump of the unmodified reference
(tests/test_host_pipeline.py::test_reader_and_normaliser_bit_exact uses the product function on the same data).
"""
import math
import warnings

import numpy as np

LOCAL_WINDOW_BP = 2000000     # mustache.py:628, 631
MIN_LOCAL_COUNT = 30          # mustache.py:657-658


def _nan_to(value, fallback):
    return fallback if math.isnan(value) else value


def normalize_sparse(x, y, v, resolution, distance_in_px):
    """In-place normalisation of `v`; returns the per-diagonal weights list the reference also returns (unused)."""
    n = max(max(x), max(y)) + 1
    weights = []
    dist = np.abs(y - x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        if (n - distance_in_px) * resolution > LOCAL_WINDOW_BP:
            box = np.ones(int(LOCAL_WINDOW_BP / resolution))
            for d in range(2 + distance_in_px):
                on_diag = dist == d
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
            for d in range(min(distance_in_px, n)):                   # mustache.py:674-675 (not 2+dpx)
                on_diag = dist == d
                g_std = _nan_to(np.std(v[on_diag]), 1)
                g_mean = _nan_to(np.mean(v[on_diag]), 0)
                v[on_diag] = (v[on_diag] - g_mean) / g_std
                np.nan_to_num(v, copy=False, nan=0, posinf=0, neginf=0)
    return weights


