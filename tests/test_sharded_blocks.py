"""The N>1 host path on CPU (gloo, world size 2): chromosomes owned by different ranks, blocks spread evenly over the
ranks (the COO of the blocks an owner cannot keep travels through all_to_all_single), every rank post-processes what it
computed, rank 0 gathers the calls -- must give exactly the single-rank result, for mustache and for diff_mustache.
The engine is replaced by a stand-in that answers with the CPU oracle (test infrastructure; the product itself has no
CPU path)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """Same surface as mustache_b200.engine.ScaleSpaceEngine for blockrun.run_batches(), computed by oracle.scalespace."""
    device = 0

    def __init__(self):
        self.tiles = {}
        self.capacity_errors = 0          # raise MB200_ERR_CAPACITY this many times before answering (retry contract)
        self.configured = []

    def set_octaves(self, octs, dedupe=True, differential=False):
        from mustache_b200 import ladder
        self.octs = list(octs)
        self.program = ladder.build_program(self.octs)

    def configure(self, n, dpx, nblocks=1, intra=True, record_fraction=-1.0):
        self.n, self.dpx, self.nblocks, self.tiles, self.diff = n, dpx, nblocks, {}, False
        self.configured.append((nblocks, record_fraction))

    def upload_coo(self, block, rows, cols, vals):
        c = np.zeros((self.n, self.n))
        c[rows, cols] = vals
        self.tiles[block] = c

    def upload_coo_batch(self, first_block, offsets, rows, cols, vals):
        for b in range(len(offsets) - 1):
            a, z = offsets[b], offsets[b + 1]
            self.upload_coo(first_block + b, rows[a:z], cols[a:z], vals[a:z])

    def run(self):
        self.diff = False

    def run_differential(self):
        self.diff = True

    def timing(self):
        return {}

    def batch_counts(self):
        return np.full(self.nblocks, 1000), np.full(self.nblocks, int(0.3 * self.n * (self.dpx - 2)))

    @staticmethod
    def _rec(st, nz_count, lut):
        f = st["p"] != 2
        r = dict(rows=st["rows"][f].astype(np.int32), cols=st["cols"][f].astype(np.int32), v=st["v"][f], p=st["p"][f],
                 score_id=st["level"][f], sigma=st["scale"][f], nz_count=nz_count, n_found=int(f.sum()))
        if "pair" in st:
            r["pair"] = st["pair"][f]
        return r

    @staticmethod
    def _empty(nz_count):
        e = np.zeros(0)
        return dict(rows=e.astype(np.int32), cols=e.astype(np.int32), v=e, p=e, score_id=e.astype(np.int32), sigma=e, pair=e,
                    nz_count=nz_count, n_found=0)

    def select_candidates(self, pt, st, candidate_fraction=-1.0):
        self.post = (pt, st)

    def _dense_candidates(self, c, nz, filled, st, pt, sthr, other=None):
        """Dense restatement of mustache.py:774-811 (and diff_mustache.py:428-500) for one map: BH, o < pt, numpy-slice
        sparsity windows, 3 x 3 neighbourhoods of the dense o / so (/ pair / v / v of the other map) matrices."""
        from oracle import postprocess as opost
        found = st["p"] != 2
        p_all = st["p"].copy()
        p_all[found] = opost.bh_statsmodels_form(st["p"][found])
        n = c.shape[0]

        def dense(mask, vals, fill=1.0):
            m = np.full((n + 2, n + 2), fill)                 # one-pixel apron of "off the tile" = 1
            m[1:-1, 1:-1][mask] = vals
            return m
        o, so = dense(nz, p_all), dense(nz, st["scale"])
        sel = found & (p_all < pt)
        x, y, sc = st["rows"][sel], st["cols"][sel], st["scale"][sel]
        keep = x != 0
        for i in range(len(x)):
            s = int(np.ceil(sc[i]))
            c1 = np.sum(nz[x[i] - s:x[i] + s + 1, y[i] - s:y[i] + s + 1]) / ((2 * s + 1) ** 2)
            s *= 2
            c2 = np.sum(nz[x[i] - s:x[i] + s + 1, y[i] - s:y[i] + s + 1]) / ((2 * s + 1) ** 2)
            if c1 < sthr or c2 < 0.6:
                keep[i] = False

        def patches(m):
            return np.array([m[a:a + 3, bb:bb + 3].ravel() for a, bb in zip(x, y)]).reshape(-1, 9)
        out = dict(rows=x.astype(np.int32), cols=y.astype(np.int32), q=p_all[sel], sigma=sc, cval=filled[x, y], keep=keep,
                   o9=patches(o), so9=patches(so), nz_count=int(nz.sum()), n_found=int(found.sum()))
        if other is not None:
            nz_o, st_o = other
            out.update(pair9=patches(dense(nz, st["pair"])), vself9=patches(dense(nz, st["v"])),
                       vother9=patches(dense(nz_o, st_o["v"])))
        return out

    @staticmethod
    def _no_candidates(nz_count, pair):
        e = np.zeros(0)
        d = dict(rows=e.astype(np.int32), cols=e.astype(np.int32), q=e, sigma=e, cval=e, keep=e.astype(bool),
                 o9=np.zeros((0, 9)), so9=np.zeros((0, 9)), nz_count=nz_count, n_found=0)
        if pair:
            d.update(pair9=np.zeros((0, 9)), vself9=np.zeros((0, 9)), vother9=np.zeros((0, 9)))
        return d

    def candidates_batch(self, pair=False):
        """What mb200_select_candidates + mb200_fetch_candidates deliver, restated densely on the oracle's output."""
        from mustache_b200.engine import EngineError
        from oracle import scalespace as osc
        if self.capacity_errors > 0:
            self.capacity_errors -= 1
            raise EngineError(-3, "block 0 produced too many records")
        pt, sthr = self.post
        out = []
        if not self.diff:
            for b in range(self.nblocks):
                c = self.tiles[b]
                res = osc.scale_space(c, self.dpx, self.octs, use_scipy=True)
                if res["skipped"]:
                    out.append(self._no_candidates(res["nz_count"], False))
                    continue
                nz, filled = osc.mask_and_fill(c, self.dpx)
                out.append(self._dense_candidates(c, nz, filled, res, pt, sthr))
            return out
        for k in range(self.nblocks // 2):
            c1, c2 = self.tiles[2 * k], self.tiles[2 * k + 1]
            res = osc.scale_space_diff(c1, c2, self.dpx, self.octs, use_scipy=True)
            if res["skipped"]:
                out += [self._no_candidates(res["nz1_count"], True), self._no_candidates(res["nz2_count"], True)]
                continue
            nz1, f1 = osc.mask_and_fill(c1, self.dpx)
            nz2, f2 = osc.mask_and_fill(c2, self.dpx)
            out.append(self._dense_candidates(c1, nz1, f1, res["map1"], pt, sthr, other=(nz2, res["map2"])))
            out.append(self._dense_candidates(c2, nz2, f2, res["map2"], pt, sthr, other=(nz1, res["map1"])))
        return out

    def records_batch(self, sort=True, pair=False, pinned=True):
        from mustache_b200.engine import EngineError
        from oracle import scalespace as osc
        if self.capacity_errors > 0:
            self.capacity_errors -= 1
            raise EngineError(-3, "block 0 produced too many records")
        out = []
        if not self.diff:
            for b in range(self.nblocks):
                res = osc.scale_space(self.tiles[b], self.dpx, self.octs, use_scipy=True)
                out.append(self._empty(res["nz_count"]) if res["skipped"] else self._rec(res, res["nz_count"], None))
            return out
        for k in range(self.nblocks // 2):
            res = osc.scale_space_diff(self.tiles[2 * k], self.tiles[2 * k + 1], self.dpx, self.octs, use_scipy=True)
            if res["skipped"]:
                out += [self._empty(res["nz1_count"]), self._empty(res["nz2_count"])]
            else:
                out += [self._rec(res["map1"], res["nz1_count"], None), self._rec(res["map2"], res["nz2_count"], None)]
        return out


def _small_geometry(n, dpx):
    chunk = max(2 * dpx, 260)
    if n <= chunk:
        return chunk, [0], [n]
    starts, ends = [0], [chunk]
    while ends[-1] < n:
        starts.append(ends[-1] - dpx)
        ends.append(starts[-1] + chunk)
    ends[-1] = n
    starts[-1] = ends[-1] - chunk
    return chunk, starts, ends


DPX = 110


def _chromosome(n, seed):
    from mustache_b200 import synth as gen
    band = gen.dense_band_tile(n, DPX, seed=seed, blob_seed=seed + 1, nblobs=40, missing=0.1)
    x, y, v = gen.band_to_coo(band, n)
    return x.astype(np.int64), y.astype(np.int64), v


CHROMS = [(900, 5), (420, 9), (640, 13)]          # 6 + 2 + 4 blocks of 260: ranks own 10 and 2, compute 6 and 6


def _patch():
    from mustache_b200 import mustache as mm
    from mustache_b200 import postprocess, tiler
    tiler.block_geometry = _small_geometry
    postprocess.MIN_MASK_FOR_BH = 1000                # the tiles are small; keep the guard but let it pass
    eng = OracleEngine()
    mm.get_engine = lambda device=None: eng
    return mm, eng


def _run_single_chromosome(rank, world):
    mm, eng = _patch()
    x, y, v = _chromosome(*CHROMS[0])
    return mm.call_blocks(x, y, v, CHROMS[0][0], DPX, [1.6, 3.2], 0.6, 0.3, verbose=False, rank=rank, world=world)


def _run_pool(rank, world):
    from mustache_b200 import sharding
    mm, eng = _patch()
    owners = sharding.chromosome_owners(len(CHROMS), world)
    preps = {c: dict(maps=[_chromosome(n, seed)], n=n) for c, (n, seed) in enumerate(CHROMS) if owners[c] == rank}
    if rank == 0:
        eng.capacity_errors = 1                       # first fetch overflows: the batch must be re-run, not lost
    out = mm.call_chromosomes(preps, len(CHROMS), DPX, [1.6, 3.2], 0.6, 0.3, verbose=False, rank=rank, world=world,
                              owners=owners)
    if rank == 0:
        assert len(eng.configured) == 2 and eng.configured[1][1] > 0.25
    return out


def _run_diff(rank, world):
    mm, eng = _patch()
    from mustache_b200 import diff_mustache as dm
    dm.get_engine = mm.get_engine
    n, seed = CHROMS[2]
    x, y, v = _chromosome(n, seed)
    rng = np.random.default_rng(77)
    v2 = v + 0.4 * rng.standard_normal(len(v))
    v2[rng.random(len(v)) < 0.05] = 0.0
    return dm.call_block_pairs((x, y, v), (x, y, v2), n, DPX, [1.6, 3.2], 0.6, 0.3, 0.5, verbose=False, rank=rank, world=world)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q.put((rank, _run_single_chromosome(rank, world), _run_pool(rank, world), _run_diff(rank, world)))
    dist.destroy_process_group()


def _same(a, b):
    key = lambda l: tuple(l[:2]) + tuple(l[4:])
    a, b = sorted(a, key=key), sorted(b, key=key)
    assert [tuple(l) for l in a] == [tuple(l) for l in b]


def test_sharded_paths_equal_single_rank():
    import torch.multiprocessing as mp
    single = _run_single_chromosome(0, 1)
    pool = _run_pool(0, 1)
    diff = _run_diff(0, 1)
    assert len(single) > 3 and sorted(pool) == [0, 1, 2] and all(len(v) > 0 for v in pool.values())
    assert len(diff) > 3 and {l[4] for l in diff} >= {1, 3}
    _same(single, pool[0])                               # chromosome 0 alone == chromosome 0 inside the pool
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=600) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] == ([], {}, [])                        # only rank 0 reports (it writes the TSV)
    _same(res[0][0], single)
    assert sorted(res[0][1]) == sorted(pool)
    for c in pool:
        _same(res[0][1][c], pool[c])
    _same(res[0][2], diff)


def test_balanced_assignment_and_owners():
    from mustache_b200 import sharding
    per = [6, 13, 19, 25, 31, 38, 44, 50]                # BASELINE config 4: 226 blocks
    for world in (1, 2, 3, 4, 8):
        owners = sharding.chromosome_owners(len(per), world, sizes=[10000 * (k + 1) for k in range(8)])
        a = sharding.balanced_assignment(per, owners, world)
        sizes = [len(a[r]) for r in range(world)]
        assert sum(sizes) == 226 and max(sizes) - min(sizes) <= 1
        flat = sorted(cb for items in a.values() for cb in items)
        assert flat == sorted((c, b) for c, nb in enumerate(per) for b in range(nb))     # a partition
        kept = sum(1 for r, items in a.items() for c, b in items if owners[c] == r)
        assert kept >= 226 // 2 or world > 2              # owners keep their own blocks first
    assert sharding.chromosome_owners(5, 2) == [0, 1, 0, 1, 0]
    words = sharding._pack_blocks([(3, 7, np.array([1, 2]), np.array([5, 9]), np.array([0.5, -1.25])), (0, 0, [], [], [])])
    back = sharding._unpack_blocks(words)
    assert back[0][:2] == (3, 7) and back[0][2].tolist() == [1, 2] and back[0][3].tolist() == [5, 9]
    assert back[0][4].tolist() == [0.5, -1.25] and back[1][:2] == (0, 0) and len(back[1][4]) == 0
