"""Launcher for the UNMODIFIED reference (ay-lab/mustache) inside the build container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (mustache_b200/) may import this.
It exists to (a) validate the restatement in oracle/ against the real reference and
(b) generate the golden vectors committed under tests/golden/ (see tests/golden/make_golden.py).
/root/reference does not exist on the GPU box, so nothing marked `gpu`, smoke() or bench.py
may call into this module.

Recipe follows SURVEY.md section 8(c):
  * empty `hicstraw` / `cooler` stub modules (imported at mustache.py:14-15, unused for text input)
  * a `statsmodels.stats.multitest.multipletests` stub restating Benjamini-Hochberg
    (statsmodels is absent from the image; reference call sites mustache.py:778, diff_mustache.py:432-433)
  * `np.Inf = np.inf` (mustache.py:234-248 use the alias removed in numpy 2)
"""
import importlib.util
import os
import runpy
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MUSTACHE_REFERENCE_ROOT", "/root/reference")

_BH_STUB = '''
import numpy as np
_HOOKS = []
def multipletests(pvals, alpha=0.05, method="fdr_bh", **kw):
    """Benjamini-Hochberg as statsmodels.stats.multitest.fdrcorrection(method="indep")."""
    assert method == "fdr_bh"
    for h in _HOOKS:
        h(pvals)
    p = np.asarray(pvals, dtype=float)
    m = p.size
    order = np.argsort(p)
    ps = np.take(p, order)
    ecdf = np.arange(1, m + 1) / float(m)
    raw = ps / ecdf
    q = np.minimum.accumulate(raw[::-1])[::-1]
    q[q > 1] = 1
    out = np.empty_like(q)
    out[order] = q
    return out <= alpha, out, None, None
'''


def _make_stub_dir():
    d = tempfile.mkdtemp(prefix="mustache_ref_stubs_")
    for name in ("hicstraw", "cooler"):
        with open(os.path.join(d, name + ".py"), "w") as f:
            f.write("# empty stub: imported but unused for text input\n")
    pkg = os.path.join(d, "statsmodels", "stats")
    os.makedirs(pkg)
    open(os.path.join(d, "statsmodels", "__init__.py"), "w").close()
    open(os.path.join(pkg, "__init__.py"), "w").close()
    with open(os.path.join(pkg, "multitest.py"), "w") as f:
        f.write(_BH_STUB)
    return d


_STUBS = None


def prepare():
    """Install the stubs on sys.path (idempotent)."""
    global _STUBS
    if _STUBS is None:
        if not os.path.isdir(REFERENCE_ROOT):
            raise RuntimeError("reference tree %s not present (only exists in the build container)" % REFERENCE_ROOT)
        _STUBS = _make_stub_dir()
        sys.path.insert(0, _STUBS)
        if not hasattr(np, "Inf"):
            np.Inf = np.inf
    return _STUBS


def load_module(which="mustache"):
    """Import mustache.py / diff_mustache.py as a module object without running main()."""
    prepare()
    import warnings
    moddir = os.path.join(REFERENCE_ROOT, "mustache")
    if which == "diff_mustache" and moddir not in sys.path:
        sys.path.insert(1, moddir)  # diff_mustache.py:16 does `from mustache import ...`
    path = os.path.join(moddir, which + ".py")
    spec = importlib.util.spec_from_file_location("_ref_" + which, path)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def bh_hooks():
    """List of callables invoked with the argument of every multipletests() call."""
    prepare()
    from statsmodels.stats import multitest
    return multitest._HOOKS


def run_cli(argv, which="mustache"):
    """Run the reference CLI unmodified: run_cli(['-f', ..., '-o', ...])."""
    prepare()
    import warnings
    moddir = os.path.join(REFERENCE_ROOT, "mustache")
    if which == "diff_mustache" and moddir not in sys.path:
        sys.path.insert(1, moddir)
    old = sys.argv
    sys.argv = [which + ".py"] + list(argv)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            runpy.run_path(os.path.join(moddir, which + ".py"), run_name="__main__")
    finally:
        sys.argv = old


if __name__ == "__main__":
    w = "mustache"
    args = sys.argv[1:]
    if args and args[0] in ("mustache", "diff_mustache"):
        w, args = args[0], args[1:]
    run_cli(args, w)
