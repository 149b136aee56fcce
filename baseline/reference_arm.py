"""Times the UNMODIFIED reference's own scale-space loop on the host cores (bench.py --impl reference / cpu_baseline).

The reference package is installed, untouched, under baseline/_ref (git-ignored; `python -m pip install --no-index
--no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`).  It needs three
things this image lacks, supplied here exactly as SURVEY.md 8(c) prescribes: empty `hicstraw` / `cooler` modules
(imported at mustache.py:14-15, unused for array input), a `statsmodels.stats.multitest.multipletests` and
`np.Inf` (removed in numpy 2).  The multipletests stand-in raises: the reference calls it at mustache.py:778, the first
statement after the scale-space loop (mustache.py:699-772), so the time from entering mustache() to that exception is
the time of the reference's hot path, measured around the reference's own code and its own scipy calls.
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_STUB = '''
class ScaleSpaceDone(Exception):
    pass
def multipletests(pvals, alpha=0.05, method="fdr_bh", **kw):
    raise ScaleSpaceDone(len(pvals))
'''


def available():
    return os.path.exists(os.path.join(REF_DIR, "mustache", "mustache.py"))


_MOD = None


def load():
    """Import the installed reference module (once per process)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    if not available():
        raise RuntimeError("baseline/_ref does not hold the reference package")
    d = tempfile.mkdtemp(prefix="mustache_ref_stubs_")
    for name in ("hicstraw", "cooler"):
        open(os.path.join(d, name + ".py"), "w").close()
    pkg = os.path.join(d, "statsmodels", "stats")
    os.makedirs(pkg)
    open(os.path.join(d, "statsmodels", "__init__.py"), "w").close()
    open(os.path.join(pkg, "__init__.py"), "w").close()
    with open(os.path.join(pkg, "multitest.py"), "w") as f:
        f.write(_STUB)
    sys.path.insert(0, d)
    sys.path.insert(0, REF_DIR)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib
        _MOD = importlib.import_module("mustache.mustache")
    return _MOD


def time_scale_space(tile, dpx, octaves, res=5000):
    """Seconds the reference spends from entering mustache() to its multipletests call, and the number of found pixels."""
    import warnings
    m = load()
    from statsmodels.stats.multitest import ScaleSpaceDone
    n = tile.shape[0]
    t0 = time.perf_counter()
    found = 0
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m.mustache(tile, "1", "1", res, [], 0, n, -1, dpx, list(octaves), 0.88, 0.1)
    except ScaleSpaceDone as done:
        found = int(done.args[0])
    return time.perf_counter() - t0, found
