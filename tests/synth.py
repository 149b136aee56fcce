"""Shared test fixtures: specs of the seeded synthetic tiles that tests/golden/make_golden.py pushed through the
unmodified reference, and writers that rebuild the reference's bundled chr21 example from the committed npz."""
import os

import numpy as np

from mustache_b200 import synth as gen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SYNTH_TILES = {
    # name: generator args, call args for mustache(c, '1','1', 5000, [], 0, N, -1, dpx, octaves, st, pt)
    "n256_o2": dict(gen=dict(n=256, dpx=100, seed=11, blob_seed=12, nblobs=12, missing=0.15),
                    dpx=100, octaves=(1.6, 3.2), st=0.7, pt=0.2),
    "n320_o4": dict(gen=dict(n=320, dpx=140, seed=21, blob_seed=22, nblobs=10, missing=0.1),
                    dpx=140, octaves=(1.6, 3.2, 6.4, 12.8), st=0.7, pt=0.2),
    "n200_full": dict(gen=dict(n=200, dpx=250, seed=31, blob_seed=32, nblobs=8, missing=0.05),
                      dpx=250, octaves=(1.6, 3.2), st=0.7, pt=0.2),   # band wider than the tile
}

SYNTH_DIFF = dict(gen=dict(n=256, dpx=100, seed=41), dpx=100, octaves=(1.6, 3.2), st=0.7, pt=0.2, pt2=0.3)


def make_tile(n, dpx, seed, blob_seed, nblobs, missing):
    band = gen.dense_band_tile(n, min(dpx, n), seed=seed, blob_seed=blob_seed, nblobs=nblobs, missing=missing)
    return gen.band_to_dense(band, n)


def make_pair(n, dpx, seed):
    a = gen.dense_band_tile(n, dpx, seed=seed, blob_seed=seed + 1, nblobs=14, missing=0.12)
    rng = np.random.default_rng(seed + 2)
    b = a + 0.35 * rng.standard_normal(a.shape)
    b[a == 0] = 0.0
    b[rng.random(a.shape) < 0.05] = 0.0          # map 2 misses a few more cells
    extra = gen.dense_band_tile(n, dpx, seed=seed + 3, blob_seed=seed + 4, nblobs=6, missing=0.0)
    noise = gen.dense_band_tile(n, dpx, seed=seed + 3, blob_seed=seed + 4, nblobs=0, missing=0.0)
    b = np.where(b != 0, b + (extra - noise), 0.0)  # blobs present only in map 2
    i = np.arange(n)[:, None]
    d = np.arange(a.shape[1])[None, :] + gen.BAND_LO
    b[(i + d) >= n] = 0.0
    return gen.band_to_dense(a, n), gen.band_to_dense(b, n)


def chr21_arrays():
    z = np.load(os.path.join(GOLDEN, "chr21_5kb_input.npz"))
    return z["bin1"].astype(np.int64), z["dist"].astype(np.int64), z["count"].astype(np.int64), z["kr"]


def write_chr21_text(outdir, counts=None, name="chr21_5kb.RAWobserved"):
    """Rebuild data/chr21_5kb.RAWobserved and .KRnorm (contacts byte for byte, biases as identical doubles; checked by make_golden.py:make_input)."""
    b1, dist, cnt, kr = chr21_arrays()
    if counts is not None:
        cnt = counts
    raw = os.path.join(outdir, name)
    p1 = b1 * 5000
    p2 = (b1 + dist) * 5000
    keep = cnt > 0
    with open(raw, "w") as f:
        f.write("".join("chr21\t%d\tchr21\t%d\t%.1f\n" % t for t in zip(p1[keep], p2[keep], cnt[keep])))
    krp = os.path.join(outdir, "chr21_5kb.KRnorm")
    with open(krp, "w") as f:
        f.write("".join("chr21\t%d\t%s\n" % (i * 5000, "NaN" if np.isnan(v) else repr(float(v))) for i, v in enumerate(kr)))
    return raw, krp


def write_chr21_thinned(outdir, keep_prob=0.6, seed=20261017):
    b1, dist, cnt, kr = chr21_arrays()
    rng = np.random.default_rng(seed)
    thin = rng.binomial(cnt, keep_prob)
    raw, _ = write_chr21_text(outdir, counts=thin, name="chr21_5kb.thinned.RAWobserved")
    return raw
