"""Seeded synthetic contact tiles (BASELINE.json configs 2-5; definitions in SURVEY.md section 8(d)).

Everything is generated in *band layout*  B[i, d - 4]  for diagonals d = j - i in [4, dpx + 1], which is the
only part of a block the reference's text reader ever populates (mustache.py:264 keeps |j-i| <= dpx+1) and
the only part the mask can select (mustache.py:699 needs j-i >= 4).  `band_to_dense` expands to the dense
N x N tile `regulator()` builds at mustache.py:923-924.
"""
import numpy as np

BAND_LO = 4  # first diagonal that can be in the mask (np.triu(c, 4), mustache.py:699)


def band_width(dpx):
    return dpx + 1 - BAND_LO + 1


def dense_band_tile(n, dpx, seed=1001, blob_seed=1002, nblobs=200, missing=0.0, dtype=np.float64):
    """Config 2: every cell with 5 <= j-i <= dpx ~ N(0,1); cells on diagonals 4 and dpx+1 non-zero;
    `nblobs` planted blobs A*exp(-r^2/2s^2), A ~ U(6,15), s in {1.5,3,6,12}; optional fraction of missing cells."""
    w = band_width(dpx)
    rng = np.random.default_rng(seed)
    band = rng.standard_normal((n, w)).astype(dtype, copy=False)
    band[band == 0] = 1.0
    if missing > 0:
        band[rng.random((n, w)) < missing] = 0.0
    brng = np.random.default_rng(blob_seed)
    sig_choices = np.array([1.5, 3.0, 6.0, 12.0])
    for _ in range(nblobs):
        s = float(brng.choice(sig_choices))
        amp = float(brng.uniform(6, 15))
        d0 = int(brng.integers(10, max(11, dpx - 10)))
        i0 = int(brng.integers(0, max(1, n - d0)))
        j0 = i0 + d0
        r = int(np.ceil(4 * s))
        ii = np.arange(max(0, i0 - r), min(n, i0 + r + 1))
        jj = np.arange(max(0, j0 - r), min(n, j0 + r + 1))
        g = amp * np.exp(-((ii[:, None] - i0) ** 2 + (jj[None, :] - j0) ** 2) / (2 * s * s))
        dd = jj[None, :] - ii[:, None]
        ok = (dd >= BAND_LO) & (dd <= dpx + 1)
        bi = np.broadcast_to(ii[:, None], dd.shape)[ok]
        bd = dd[ok] - BAND_LO
        keep = band[bi, bd] != 0
        band[bi[keep], bd[keep]] += g[ok][keep]
    # clip to the tile: (i, i+d) must have i+d < n
    i = np.arange(n)[:, None]
    d = np.arange(w)[None, :] + BAND_LO
    band[(i + d) >= n] = 0.0
    return band


def band_to_dense(band, n=None):
    n = band.shape[0] if n is None else n
    w = band.shape[1]
    c = np.zeros((n, n), dtype=band.dtype)
    i = np.arange(band.shape[0])[:, None]
    j = i + np.arange(w)[None, :] + BAND_LO
    ok = j < n
    c[np.broadcast_to(i, j.shape)[ok], j[ok]] = band[ok]
    return c


def band_to_coo(band, n=None):
    n = band.shape[0] if n is None else n
    i, k = np.nonzero(band)
    j = i + k + BAND_LO
    ok = j < n
    return i[ok].astype(np.int32), j[ok].astype(np.int32), band[i[ok], k[ok]]


def poisson_chromosome(n, dpx, lam_scale=18.0, seed=3000, nloops=None, loop_seed=None, loop_boost=8.0):
    """Configs 3/4: raw counts ~ Poisson(lam(d)), lam(d) = lam_scale/(d+1) for d >= 1 (30 at d = 0), bias == 1,
    planted loops adding Poisson(loop_boost) in a 3x3 patch.  Returns upper-triangular COO (x, y, count) with |y-x| <= dpx+1,
    i.e. what read_pd() would return for a 3/5-column text file of these counts (mustache.py:254-297)."""
    rng = np.random.default_rng(seed)
    xs, ys, vs = [], [], []
    for d in range(0, dpx + 2):
        lam = 30.0 if d == 0 else lam_scale / (d + 1)
        m = n - d
        if m <= 0:
            break
        cnt = rng.poisson(lam, m)
        nzi = np.nonzero(cnt)[0]
        xs.append(nzi)
        ys.append(nzi + d)
        vs.append(cnt[nzi])
    x = np.concatenate(xs).astype(np.int64)
    y = np.concatenate(ys).astype(np.int64)
    v = np.concatenate(vs).astype(np.float64)
    if nloops:
        lrng = np.random.default_rng(seed + 1 if loop_seed is None else loop_seed)
        import scipy.sparse as sp
        a = sp.coo_matrix((v, (x, y)), shape=(n, n)).tolil()
        for _ in range(nloops):
            d0 = int(lrng.integers(12, dpx - 4))
            i0 = int(lrng.integers(2, n - d0 - 2))
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    a[i0 + di, i0 + d0 + dj] += lrng.poisson(loop_boost)
        a = a.tocoo()
        keep = a.data > 0
        x, y, v = a.row[keep].astype(np.int64), a.col[keep].astype(np.int64), a.data[keep].astype(np.float64)
    return x, y, v


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs 3 / 4 / 5: whole synthetic chromosomes as raw count maps
# ------------------------------------------------------------------------------------------------------------------
def _plant_loops(band, n, dpx, nloops, rng, boost, sign=+1, dmax=None):
    """Adds Poisson(boost) counts in a 3x3 patch around `nloops` random centres at distances [12, dpx - 4), or
    [8, dmax) when dmax is given: the reference's sparsity filter (mustache.py:800-811) only keeps calls whose
    (4s+1)^2 neighbourhood is >= 60 % non-zero, which at these depths means close to the diagonal."""
    d0 = rng.integers(12, dpx - 4, size=nloops) if dmax is None else rng.integers(8, min(dmax, dpx - 4), size=nloops)
    i0 = (rng.random(nloops) * (n - d0 - 4)).astype(np.int64) + 2
    inc = rng.poisson(boost, size=(nloops, 3, 3))
    for a, di in enumerate((-1, 0, 1)):
        for b, dj in enumerate((-1, 0, 1)):
            np.add.at(band, (i0 + di, d0 + dj - di), sign * inc[:, a, b])
    return i0, d0


def poisson_band(n, dpx, lam_scale, seed, chunk_rows=2048):
    """Raw counts in band layout [n][dpx + 2] (column = j - i, 0 .. dpx+1): Poisson(30) on the diagonal,
    Poisson(lam_scale / (d + 1)) at distance d >= 1; cells past the chromosome end are 0."""
    w = dpx + 2
    lam = np.empty(w)
    lam[0] = 30.0
    lam[1:] = lam_scale / (np.arange(1, w) + 1.0)
    rng = np.random.default_rng(seed)
    band = np.empty((n, w), dtype=np.int32)
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        band[r0:r1] = rng.poisson(lam[None, :], size=(r1 - r0, w))
    return band


def _clip_band(band, n):
    w = band.shape[1]
    tail = min(n, w)
    i = np.arange(n - tail, n)[:, None]
    band[n - tail:][(i + np.arange(w)[None, :]) >= n] = 0
    return band


def band_counts_to_coo(band):
    """Row-major upper-triangular COO (x, y, count) of a count band: the line order of the text files we write."""
    i, k = np.nonzero(band)
    return i.astype(np.int64), (i + k).astype(np.int64), band[i, k].astype(np.int64)


def synthetic_chromosome(n, dpx, lam_scale, seed, nloops, loop_seed=None, loop_boost=8.0, loop_dmax=None):
    """One synthetic chromosome of configs 3 / 4: Poisson background + `nloops` planted 3x3 loops; COO of raw counts."""
    band = poisson_band(n, dpx, lam_scale, seed)
    if nloops:
        _plant_loops(band, n, dpx, nloops, np.random.default_rng(seed + 1 if loop_seed is None else loop_seed), loop_boost,
                     dmax=loop_dmax)
    return band_counts_to_coo(_clip_band(band, n))


# name -> (n bins, resolution, dpx, lam_scale, background seed, loops, loop seed)
CONFIG3 = dict(n=50000, res=1000, dpx=2000, lam_scale=4.0, seed=2001, nloops=2000, loop_seed=2002)
# same geometry as config 3 (1 kb, N 4000, dpx 2000) but dense enough near the diagonal for loops to survive the
# reference's sparsity filter (mustache.py:800-811 needs >= 60 % non-zero pixels in a (4s+1)^2 window): config 3 as
# SURVEY defines it yields 0 loops in the reference
CONFIG3D = dict(n=12000, res=1000, dpx=2000, lam_scale=60.0, seed=2101, nloops=300, loop_seed=2102, loop_dmax=60)
CONFIG4 = {"s%d" % (k + 1): dict(n=10000 * (k + 1), res=5000, dpx=400, lam_scale=18.0, seed=3000 + k, nloops=40 * (k + 1),
                                 loop_seed=3100 + k, loop_dmax=24, loop_boost=25.0) for k in range(8)}
CONFIG5 = dict(n=20000, res=5000, dpx=400, lam_scale=18.0, seed=4001, loop_seed=4002, thin_seed=4003, nloops=300, ndelete=100,
               nadd=100, keep_prob=0.6, loop_dmax=24, loop_boost=25.0)


def config5_maps(n=20000, dpx=400, lam_scale=18.0, seed=4001, loop_seed=4002, thin_seed=4003, nloops=300, ndelete=100,
                 nadd=100, keep_prob=0.6, loop_boost=8.0, loop_dmax=None, **_):
    """Config 5: map A = background + `nloops` loops; map B = binomial(keep_prob) thinning of map A without its last
    `ndelete` loops, plus `nadd` loops of its own.  Returns two COO triples of raw counts."""
    base = poisson_band(n, dpx, lam_scale, seed)
    lrng = np.random.default_rng(loop_seed)
    common = base.copy()
    _plant_loops(common, n, dpx, nloops - ndelete, lrng, loop_boost, dmax=loop_dmax)
    a = common.copy()
    _plant_loops(a, n, dpx, ndelete, lrng, loop_boost, dmax=loop_dmax)
    trng = np.random.default_rng(thin_seed)
    b = trng.binomial(common, keep_prob).astype(np.int32)
    _plant_loops(b, n, dpx, nadd, trng, loop_boost, dmax=loop_dmax)
    return band_counts_to_coo(_clip_band(a, n)), band_counts_to_coo(_clip_band(b, n))


def write_contact_text(path, chrom, x, y, counts, res, mode="w"):
    """5-column contact text (chr pos chr pos count) in the style of the reference's bundled data/chr21_5kb.RAWobserved:
    counts written as '3.0', so that read_pd (mustache.py:254-297) yields float64 values as it does for that file."""
    name = str(chrom)
    with open(path, mode) as f:
        step = 1 << 20
        for s in range(0, len(x), step):
            xs, ys, cs = x[s:s + step] * res, y[s:s + step] * res, counts[s:s + step]
            f.write("".join([name + "\t%d\t" % a + name + "\t%d\t%d.0\n" % (b, c) for a, b, c in zip(xs.tolist(), ys.tolist(), cs.tolist())]))
    return path
