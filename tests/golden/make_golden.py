"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Usage:  python tests/golden/make_golden.py [what ...]
  what in {input, cli, blocks, synth, diff, diffcli}; default = all.

Outputs (all committed):
  chr21_5kb_input.npz      compact copy of the reference's bundled example input (data/chr21_5kb.*), the only
                           runnable example the reference ships (README.md:49-51) -- DATA, not source.
  chr21_loops.tsv          G1: sorted output of `mustache -f chr21_5kb.RAWobserved -b chr21_5kb.KRnorm -ch 21
                           -r 5kb -pt 0.1 -st 0.8`
  chr21_blocks.npz         G2: per block, what mustache() holds at the multipletests call (mustache.py:778):
                           found (row, col, v, scale, p_raw), mask size; plus the loops each block returned and a
                           digest of the normalised COO that regulator() fed the tiler
  synth_*.npz              G3: same dumps for seeded synthetic tiles pushed through mustache() directly
  diff_synth.npz           diff_mustache() dumps for a seeded synthetic pair
  chr21_diff_*.tsv         diff CLI outputs for chr21 vs its binomial(0.6) thinning (seed 20261017)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import run_reference as rr  # noqa: E402
from tests import synth  # noqa: E402

REF_DATA = os.path.join(rr.REFERENCE_ROOT, "data")


def make_input():
    import pandas as pd
    df = pd.read_csv(os.path.join(REF_DATA, "chr21_5kb.RAWobserved"), sep="\t", header=None)
    kr = pd.read_csv(os.path.join(REF_DATA, "chr21_5kb.KRnorm"), sep="\t", header=None)
    assert (df[1] % 5000 == 0).all() and (df[3] % 5000 == 0).all() and (df[1] <= df[3]).all()
    assert (df[4] == np.round(df[4])).all()
    # read_bias() parses with Python float() (mustache.py:231) -- exact, unlike pandas' default fast parser
    kr_lines = [l.split("\t") for l in open(os.path.join(REF_DATA, "chr21_5kb.KRnorm")).read().splitlines()]
    assert all(int(l[1]) == 5000 * i for i, l in enumerate(kr_lines))
    kr_exact = np.array([float(l[2]) for l in kr_lines], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "chr21_5kb_input.npz"),
                        bin1=(df[1] // 5000).astype(np.int16),
                        dist=((df[3] - df[1]) // 5000).astype(np.int16),
                        count=df[4].astype(np.int32),
                        kr=kr_exact)
    # the contact file must be reproducible byte for byte, the bias file value for value from the npz (tests/synth.py:write_chr21_text)
    tmp = "/tmp/_golden_chk"
    os.makedirs(tmp, exist_ok=True)
    raw, krp = synth.write_chr21_text(tmp)
    assert open(raw, "rb").read() == open(os.path.join(REF_DATA, "chr21_5kb.RAWobserved"), "rb").read()
    # the bias file carries 17 significant digits; the rebuilt one uses the shortest repr of the same doubles
    k2 = np.array([float(l.split("\t")[2]) for l in open(krp).read().splitlines()])
    assert np.array_equal(k2, kr_exact, equal_nan=True)
    print("input ok")


def make_cli():
    out = "/tmp/_golden_chr21.tsv"
    rr.run_cli(["-f", os.path.join(REF_DATA, "chr21_5kb.RAWobserved"), "-b", os.path.join(REF_DATA, "chr21_5kb.KRnorm"),
                "-ch", "21", "-r", "5kb", "-pt", "0.1", "-st", "0.8", "-p", "8", "-o", out])
    lines = open(out).read().splitlines()
    body = sorted(lines[1:], key=lambda s: (int(s.split("\t")[1]), int(s.split("\t")[4])))
    with open(os.path.join(HERE, "chr21_loops.tsv"), "w") as f:
        f.write("\n".join([lines[0]] + body) + "\n")
    print("cli:", len(body), "loops; sorted md5", hashlib.md5(("\n".join(sorted(lines)) + "\n").encode()).hexdigest())


class Dump:
    """Capture mustache()'s locals at the multipletests call (its argument is pAll[pFound], mustache.py:778)."""

    def __init__(self):
        self.items = []

    def __call__(self, pvals):
        fr = sys._getframe(2).f_locals
        if "pAll" in fr:                      # mustache()
            nz, found = fr["nz"], fr["pFound"]
            rows, cols = np.nonzero(nz)
            self.items.append(dict(rows=rows[found].astype(np.int32), cols=cols[found].astype(np.int32),
                                   v=fr["vAll"][found].copy(), scale=fr["Scales"][found].copy(),
                                   p=fr["pAll"][found].copy(), nz_count=int(nz.sum())))
        else:                                 # diff_mustache(): first call map 1, second call map 2
            which = "1" if len([i for i in self.items if "pair" in i]) % 2 == 0 else "2"
            nz, found = fr["nz" + which], fr["pFound" + which]
            rows, cols = np.nonzero(nz)
            self.items.append(dict(rows=rows[found].astype(np.int32), cols=cols[found].astype(np.int32),
                                   v=fr["vAll" + which][found].copy(), scale=fr["Scales" + which][found].copy(),
                                   p=fr["pAll" + which][found].copy(), pair=fr["pPair" + which][found].copy(),
                                   nz_count=int(nz.sum())))


def _pack(prefix, d, out):
    for k, val in d.items():
        out[prefix + k] = np.asarray(val)


def _loops_arr(loops):
    return np.array([[float(a) for a in l] for l in loops], dtype=np.float64).reshape(-1, 4 if not loops else len(loops[0]))


def make_blocks():
    """Replay regulator()'s tiling (mustache.py:892-924) on chr21 with the reference's own functions."""
    import math
    m = rr.load_module("mustache")
    hooks = rr.bh_hooks()
    res, dist_bp = 5000, 2000000
    x, y, v = m.read_pd(os.path.join(REF_DATA, "chr21_5kb.RAWobserved"), dist_bp,
                        os.path.join(REF_DATA, "chr21_5kb.KRnorm"), "21", res)
    x, y = np.asarray(x), np.asarray(y)
    dpx = int(math.ceil(dist_bp // res))
    n = max(max(x), max(y)) + 1
    m.normalize_sparse(x, y, v, res, dpx)
    out = dict(n=n, dpx=dpx, nnz=len(v),
               coo_digest=np.frombuffer(hashlib.sha256(np.ascontiguousarray(x, np.int64).tobytes()
                                                       + np.ascontiguousarray(y, np.int64).tobytes()
                                                       + np.ascontiguousarray(v, np.float64).tobytes()).digest(), np.uint8),
               v_sum=np.float64(v.sum()))
    chunk = max(2 * dpx, 2000)
    start, end = [0], [chunk]
    while end[-1] < n:
        start.append(end[-1] - dpx)
        end.append(start[-1] + chunk)
    end[-1] = n
    start[-1] = end[-1] - chunk
    out["start"], out["end"] = np.array(start), np.array(end)
    for b in range(len(start)):
        sel = (x >= start[b]) & (x < end[b]) & (y >= start[b]) & (y < end[b])
        cc = np.zeros((chunk, chunk))
        cc[x[sel] - start[b], y[sel] - start[b]] = v[sel]
        d = Dump()
        hooks.append(d)
        loops = m.mustache(cc, "21", "21", res, [], start[b], end[b], -1, dpx, [1.6, 3.2], 0.8, 0.1)
        hooks.remove(d)
        out["b%d_loops" % b] = _loops_arr(loops)
        if d.items:
            _pack("b%d_" % b, d.items[0], out)
        else:
            out["b%d_nz_count" % b] = int(((cc != 0) & (np.triu(np.ones_like(cc), 4) > 0)).sum())
        print("block", b, "loops", len(loops), "found", len(d.items[0]["p"]) if d.items else None)
    np.savez_compressed(os.path.join(HERE, "chr21_blocks.npz"), **out)


def make_synth():
    m = rr.load_module("mustache")
    hooks = rr.bh_hooks()
    for name, spec in synth.SYNTH_TILES.items():
        cc = synth.make_tile(**spec["gen"])
        d = Dump()
        hooks.append(d)
        loops = m.mustache(cc.copy(), "1", "1", 5000, [], 0, cc.shape[0], -1, spec["dpx"], list(spec["octaves"]),
                           spec["st"], spec["pt"])
        hooks.remove(d)
        out = dict(loops=_loops_arr(loops), tile_digest=np.frombuffer(hashlib.sha256(cc.tobytes()).digest(), np.uint8))
        _pack("", d.items[0], out)
        np.savez_compressed(os.path.join(HERE, "synth_%s.npz" % name), **out)
        print(name, "mask", d.items[0]["nz_count"], "found", len(d.items[0]["p"]), "loops", len(loops))


def make_diff():
    dm = rr.load_module("diff_mustache")
    hooks = rr.bh_hooks()
    spec = synth.SYNTH_DIFF
    c1, c2 = synth.make_pair(**spec["gen"])
    d = Dump()
    hooks.append(d)
    r = dm.diff_mustache(c1.copy(), c2.copy(), "1", "1", 5000, 0, c1.shape[0], -1, spec["dpx"], list(spec["octaves"]),
                         spec["st"], spec["pt"], spec["pt2"])
    hooks.remove(d)
    out = dict(digest=np.frombuffer(hashlib.sha256(c1.tobytes() + c2.tobytes()).digest(), np.uint8))
    for nm, loops in zip(("loops1", "diff1", "loops2", "diff2"), r):
        out[nm] = _loops_arr(loops)
    _pack("m1_", d.items[0], out)
    _pack("m2_", d.items[1], out)
    np.savez_compressed(os.path.join(HERE, "diff_synth.npz"), **out)
    print("diff: found", len(d.items[0]["p"]), len(d.items[1]["p"]), "loops", [len(a) for a in r])


def make_diffcli():
    tmp = "/tmp/_golden_diff"
    os.makedirs(tmp, exist_ok=True)
    raw, kr = synth.write_chr21_text(tmp)
    thin = synth.write_chr21_thinned(tmp)
    outp = os.path.join(tmp, "out")
    rr.run_cli(["-f1", raw, "-f2", thin, "-b1", kr, "-b2", kr, "-ch", "21", "-r", "5kb", "-pt", "0.05", "-pt2", "0.1",
                "-st", "0.8", "-p", "8", "-o", outp], which="diff_mustache")
    for suf in ("loop1", "loop2", "diffloop1", "diffloop2"):
        lines = open(outp + "." + suf).read().splitlines()
        body = sorted(lines[1:], key=lambda s: (int(s.split("\t")[1]), int(s.split("\t")[4])))
        with open(os.path.join(HERE, "chr21_diff_%s.tsv" % suf), "w") as f:
            f.write("\n".join([lines[0]] + body) + "\n")
        print(suf, len(body))


def _sorted_body(path):
    lines = open(path).read().splitlines()
    return lines[0], sorted(lines[1:], key=lambda s: (s.split("\t")[0], int(s.split("\t")[1]), int(s.split("\t")[4])))


def _res_arg(res):
    return "%dkb" % (res // 1000)


def make_cfg3():
    """BASELINE configs[2]: synthetic 50k-bin chromosome at 1 kb through the reference CLI, plus the
    mustache() dumps of two of its 4000 x 4000 blocks (replaying regulator()'s tiling in-process)."""
    import math
    import time
    from mustache_b200 import synth as gen
    spec = gen.CONFIG3
    tmp = "/tmp/_golden_cfg3"
    os.makedirs(tmp, exist_ok=True)
    x, y, c = gen.synthetic_chromosome(**{k: v for k, v in spec.items() if k != "res"})
    path = gen.write_contact_text(os.path.join(tmp, "cfg3.txt"), "chrS", x, y, c, spec["res"])
    out = os.path.join(tmp, "out.tsv")
    t0 = time.time()
    rr.run_cli(["-f", path, "-ch", "chrS", "-r", _res_arg(spec["res"]), "-pt", "0.1", "-st", "0.8", "-p", "8", "-o", out])
    print("cfg3 reference CLI: %.1f s" % (time.time() - t0))
    head, body = _sorted_body(out)
    with open(os.path.join(HERE, "cfg3_loops.tsv"), "w") as f:
        f.write("\n".join([head] + body) + "\n")
    print("cfg3:", len(body), "loops")
    # block dumps
    m = rr.load_module("mustache")
    hooks = rr.bh_hooks()
    res, dpx = spec["res"], spec["dpx"]
    xx, yy, vv = m.read_pd(path, 2000000, False, "chrS", res)
    xx, yy = np.asarray(xx), np.asarray(yy)
    m.normalize_sparse(xx, yy, vv, res, dpx)
    n = max(max(xx), max(yy)) + 1
    chunk = max(2 * dpx, 2000)
    outz = dict(n=n, dpx=dpx, nnz=len(vv), v_sum=np.float64(vv.sum()))
    for b in CFG3_DUMP_BLOCKS:
        s0 = b * (chunk - dpx)
        sel = (xx >= s0) & (xx < s0 + chunk) & (yy >= s0) & (yy < s0 + chunk)
        cc = np.zeros((chunk, chunk))
        cc[xx[sel] - s0, yy[sel] - s0] = vv[sel]
        d = Dump()
        hooks.append(d)
        loops = m.mustache(cc, "chrS", "chrS", res, [], s0, s0 + chunk, -1, dpx, [1.6, 3.2], 0.8, 0.1)
        hooks.remove(d)
        outz["b%d_loops" % b] = _loops_arr(loops)
        _pack("b%d_" % b, d.items[0], outz)
        print("cfg3 block", b, "mask", d.items[0]["nz_count"], "found", len(d.items[0]["p"]), "loops", len(loops))
    np.savez_compressed(os.path.join(HERE, "cfg3_blocks.npz"), **outz)


CFG3_DUMP_BLOCKS = (0, 11)


def make_cfg3d():
    """Denser 1 kb chromosome (same block geometry as config 3): the reference CLI finds loops here."""
    import time
    from mustache_b200 import synth as gen
    spec = gen.CONFIG3D
    tmp = "/tmp/_golden_cfg3d"
    os.makedirs(tmp, exist_ok=True)
    x, y, c = gen.synthetic_chromosome(**{k: v for k, v in spec.items() if k != "res"})
    path = gen.write_contact_text(os.path.join(tmp, "cfg3d.txt"), "chrT", x, y, c, spec["res"])
    out = os.path.join(tmp, "out.tsv")
    t0 = time.time()
    rr.run_cli(["-f", path, "-ch", "chrT", "-r", _res_arg(spec["res"]), "-pt", "0.1", "-st", "0.8", "-p", "8", "-o", out])
    head, body = _sorted_body(out)
    with open(os.path.join(HERE, "cfg3d_loops.tsv"), "w") as f:
        f.write("\n".join([head] + body) + "\n")
    print("cfg3d:", len(body), "loops, %.1f s" % (time.time() - t0))


def make_cfg4():
    """BASELINE configs[3]: the 8 synthetic chromosomes (10k..80k bins at 5 kb), reference CLI per chromosome."""
    import time
    from mustache_b200 import synth as gen
    tmp = "/tmp/_golden_cfg4"
    os.makedirs(tmp, exist_ok=True)
    rows, head = [], None
    for name, spec in gen.CONFIG4.items():
        x, y, c = gen.synthetic_chromosome(**{k: v for k, v in spec.items() if k != "res"})
        path = gen.write_contact_text(os.path.join(tmp, name + ".txt"), name, x, y, c, spec["res"])
        out = os.path.join(tmp, name + ".tsv")
        t0 = time.time()
        rr.run_cli(["-f", path, "-ch", name, "-r", _res_arg(spec["res"]), "-pt", "0.1", "-st", "0.8", "-p", "8", "-o", out])
        head, body = _sorted_body(out)
        rows += body
        print("cfg4", name, len(body), "loops, %.1f s" % (time.time() - t0), flush=True)
    with open(os.path.join(HERE, "cfg4_loops.tsv"), "w") as f:
        f.write("\n".join([head] + rows) + "\n")


def make_cfg5():
    """BASELINE configs[4]: differential CLI on two synthetic 20k-bin maps, -pt 0.05 -pt2 0.1."""
    from mustache_b200 import synth as gen
    spec = gen.CONFIG5
    tmp = "/tmp/_golden_cfg5"
    os.makedirs(tmp, exist_ok=True)
    A, B = gen.config5_maps(**spec)
    fa = gen.write_contact_text(os.path.join(tmp, "mapA.txt"), "chrD", *A, spec["res"])
    fb = gen.write_contact_text(os.path.join(tmp, "mapB.txt"), "chrD", *B, spec["res"])
    outp = os.path.join(tmp, "out")
    rr.run_cli(["-f1", fa, "-f2", fb, "-ch", "chrD", "-r", _res_arg(spec["res"]), "-pt", "0.05", "-pt2", "0.1", "-st", "0.8",
                "-p", "8", "-o", outp], which="diff_mustache")
    for suf in ("loop1", "loop2", "diffloop1", "diffloop2"):
        head, body = _sorted_body(outp + "." + suf)
        with open(os.path.join(HERE, "cfg5_%s.tsv" % suf), "w") as f:
            f.write("\n".join([head] + body) + "\n")
        print("cfg5", suf, len(body))


class _FitSpy:
    """Stands in for `expon` inside the reference module: records every fit() result, forwards everything else."""

    def __init__(self, real):
        self._real, self.fits = real, []

    def fit(self, *a, **k):
        r = self._real.fit(*a, **k)
        self.fits.append(r)
        return r

    def __getattr__(self, name):
        return getattr(self._real, name)


def make_cfg2():
    """BASELINE configs[1] at full size: the 10k x 10k dense band, 4 octaves, through the reference's mustache() once
    (about 10 minutes and 13 GB on one core).  Commits the 36 exponential fits, a digest of the sorted records and a
    sample of the p-values -- the records themselves (1.3 M) are too large for a fixture."""
    import time
    from mustache_b200 import synth as gen
    m = rr.load_module("mustache")
    hooks = rr.bh_hooks()
    n, dpx, octs = 10000, 5000, [1.6, 3.2, 6.4, 12.8]
    cc = gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=1001, blob_seed=1002, nblobs=200), n)
    spy = _FitSpy(m.expon)
    m.expon = spy
    d = Dump()
    hooks.append(d)
    t0 = time.time()
    loops = m.mustache(cc, "1", "1", 2000, [], 0, n, -1, dpx, octs, 0.88, 0.1)
    hooks.remove(d)
    m.expon = spy._real
    print("cfg2 reference mustache(): %.1f s" % (time.time() - t0))
    it = d.items[0]
    h = hashlib.sha256()
    for k in ("rows", "cols", "v", "scale"):
        h.update(np.ascontiguousarray(it[k]).tobytes())
    step = 997
    np.savez_compressed(os.path.join(HERE, "cfg2_tile.npz"), fits=np.array(spy.fits, dtype=np.float64),
                        n_found=len(it["p"]), nz_count=it["nz_count"], digest=np.frombuffer(h.digest(), np.uint8),
                        sample_step=step, sample_rows=it["rows"][::step], sample_cols=it["cols"][::step],
                        sample_v=it["v"][::step], sample_scale=it["scale"][::step], sample_p=it["p"][::step],
                        p_sum=np.float64(it["p"].sum()), loops=_loops_arr(loops))
    print("cfg2: mask", it["nz_count"], "found", len(it["p"]), "loops", len(loops), "fits", len(spy.fits))


if __name__ == "__main__":
    what = sys.argv[1:] or ["input", "cli", "blocks", "synth", "diff", "diffcli"]
    for w in what:
        globals()["make_" + w]()
