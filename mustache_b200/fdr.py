"""Benjamini-Hochberg, as statsmodels.stats.multitest.multipletests(method='fdr_bh') computes it
(reference call sites mustache.py:778, diff_mustache.py:432-433; statsmodels' fdrcorrection(method='indep'):
sort, p/ecdf with ecdf = arange(1, m+1)/m, reverse cumulative minimum, clip to 1, unsort)."""
import numpy as np


def fdr_bh(pvals):
    p = np.asarray(pvals, dtype=np.float64)
    m = p.size
    if m == 0:
        return p.copy()
    order = np.argsort(p)
    ranked = np.take(p, order)
    ranked = ranked / (np.arange(1, m + 1) / float(m))
    ranked = np.minimum.accumulate(ranked[::-1])[::-1]
    ranked[ranked > 1] = 1
    q = np.empty_like(ranked)
    q[order] = ranked
    return q
