"""The N>1 host path on CPU (gloo, world size 2): blocks sharded round-robin over ranks, records all-gathered, rank 0
post-processes -- must give exactly the single-rank result.  The engine is replaced by a stand-in that answers with the
CPU oracle (test infrastructure; the product itself has no CPU path)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """Same surface as mustache_b200.engine.ScaleSpaceEngine for call_blocks(), computed by oracle.scalespace."""
    device = 0

    def __init__(self):
        self.tiles = {}

    def set_octaves(self, octs, dedupe=True, differential=False):
        from mustache_b200 import ladder
        self.octs = list(octs)
        self.program = ladder.build_program(self.octs)

    def _sigma_lut(self):
        lut = np.zeros(256)
        for k, sg in self.program.sigma_of_id.items():
            lut[k] = sg
        return lut

    def configure(self, n, dpx, nblocks=1, intra=True, record_fraction=-1.0):
        self.n, self.dpx, self.tiles = n, dpx, {}

    def upload_coo(self, block, rows, cols, vals):
        c = np.zeros((self.n, self.n))
        c[rows, cols] = vals
        self.tiles[block] = c

    def run(self):
        pass

    def timing(self):
        return {}

    def records(self, block, sort=True, pair=False, pinned=False):
        from oracle import scalespace as osc
        res = osc.scale_space(self.tiles[block], self.dpx, self.octs, use_scipy=True)
        if res["skipped"]:
            e = np.zeros(0)
            return dict(rows=e.astype(np.int32), cols=e.astype(np.int32), v=e, p=e, score_id=e.astype(np.int32), sigma=e,
                        nz_count=res["nz_count"], n_found=0)
        f = res["p"] != 2
        return dict(rows=res["rows"][f].astype(np.int32), cols=res["cols"][f].astype(np.int32), v=res["v"][f], p=res["p"][f],
                    score_id=res["level"][f], sigma=res["scale"][f], nz_count=res["nz_count"], n_found=int(f.sum()))


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


def _problem():
    from mustache_b200 import synth as gen
    n, dpx = 900, 110
    band = gen.dense_band_tile(n, dpx, seed=5, blob_seed=6, nblobs=40, missing=0.1)
    x, y, v = gen.band_to_coo(band, n)
    return x.astype(np.int64), y.astype(np.int64), v, n, dpx


def _run(rank, world):
    from mustache_b200 import mustache as mm
    from mustache_b200 import tiler
    tiler.block_geometry = _small_geometry
    eng = OracleEngine()
    mm.get_engine = lambda device=None: eng
    x, y, v, n, dpx = _problem()
    return mm.call_blocks(x, y, v, n, dpx, [1.6, 3.2], 0.6, 0.3, verbose=False, rank=rank, world=world)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q.put((rank, _run(rank, world)))
    dist.destroy_process_group()


def test_sharded_call_blocks_equals_single_rank():
    import torch.multiprocessing as mp
    single = _run(0, 1)
    assert len(single) > 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] == []                                  # only rank 0 reports (it writes the TSV)
    key = lambda l: (int(l[0]), int(l[1]))
    a, b = sorted(res[0], key=key), sorted(single, key=key)
    assert [key(l) for l in a] == [key(l) for l in b]
    assert [(l[2], l[3]) for l in a] == [(l[2], l[3]) for l in b]
