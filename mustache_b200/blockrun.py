"""The block pool of a run: tiles every chromosome, spreads the blocks over the ranks, pushes each rank's share through its
engine in batches and post-processes them where they were computed.

Shared by mustache.py (one map per block) and diff_mustache.py (two maps per block).  Replaces the reference's
per-chromosome process fan-out and its Manager().list() result sink (mustache.py:913-937, 945-960;
diff_mustache.py:654-717).  See sharding.py for the multi-GPU scheme.
"""
import os
import time

import numpy as np

from . import sharding, tiler

# tile slots (double buffered) of one engine batch stay below this many bytes; records and scratch come on top
MAX_BATCH_TILE_BYTES = int(float(os.environ.get("MUSTACHE_BATCH_GB", "12")) * (1 << 30))


class BlockTask:
    """One block of one chromosome: `maps` holds, per contact map (1 for mustache, 2 for diff_mustache), the mask pixels
    (rows, cols, vals) of the tile regulator() would build at mustache.py:919-924."""
    __slots__ = ("chrom", "block", "maps")

    def __init__(self, chrom, block, maps):
        self.chrom, self.block, self.maps = chrom, block, maps


def collective_device(eng):
    import torch
    return torch.device("cuda", eng.device) if torch.cuda.is_available() else torch.device("cpu")


def build_tasks(preps, n_chrom, dpx, nmaps, rank, world, eng, owners=None):
    """preps: {chromosome index: [(x, y, v) per map] + n} for the chromosomes THIS rank read and normalised.
    Returns (tasks of this rank, geometry {chrom: (chunk, starts, ends)})."""
    owners = owners if owners is not None else sharding.chromosome_owners(n_chrom, world)
    mine = {c: int(p["n"]) for c, p in preps.items()}
    lengths = {}
    for part in sharding.all_gather_meta(mine, world):
        lengths.update(part)
    geom = {c: tiler.block_geometry(lengths[c], dpx) for c in range(n_chrom) if lengths.get(c, 0) > 0}
    per_chrom = [len(geom[c][1]) if c in geom else 0 for c in range(n_chrom)]
    assign = sharding.balanced_assignment(per_chrom, owners, world)
    where = {cb: r for r, items in assign.items() for cb in items}
    tasks, send = [], {}
    for c, p in preps.items():
        if c not in geom:
            continue
        chunk, starts, ends = geom[c]
        slicers = [tiler.BlockSlicer(*m) for m in p["maps"]]
        for b in range(len(starts)):
            maps = [tiler.block_mask_pixels(*s.block(starts[b], ends[b]), chunk) for s in slicers]
            dst = where[(c, b)]
            if dst == rank:
                tasks.append(BlockTask(c, b, maps))
            else:
                for w, m in enumerate(maps):
                    send.setdefault(dst, []).append((c, b * nmaps + w, m[0], m[1], m[2]))
    if world > 1:
        got = {}
        for c, bw, rows, cols, vals in sharding.exchange_blocks(send, rank, world, collective_device(eng)):
            got.setdefault((c, bw // nmaps), {})[bw % nmaps] = (rows, cols, vals)
        for (c, b), maps in got.items():
            tasks.append(BlockTask(c, b, [maps[w] for w in range(nmaps)]))
    tasks.sort(key=lambda t: (t.chrom, t.block))
    assert sorted((t.chrom, t.block) for t in tasks) == sorted(assign[rank])
    return tasks, geom


def concat_coo(maps):
    """[(rows, cols, vals)] -> (offsets, rows int32, cols int32, vals float64) for ScaleSpaceEngine.upload_coo_batch."""
    offsets = np.zeros(len(maps) + 1, np.int64)
    np.cumsum([len(m[2]) for m in maps], out=offsets[1:])
    tot = int(offsets[-1])
    rows, cols, vals = np.empty(tot, np.int32), np.empty(tot, np.int32), np.empty(tot, np.float64)
    for k, m in enumerate(maps):
        a, z = offsets[k], offsets[k + 1]
        rows[a:z], cols[a:z], vals[a:z] = m[0], m[1], m[2]
    return offsets, rows, cols, vals


def run_batches(eng, tasks, chunk, dpx, differential=False, verbose=False, timings=None, totals=None, select=None):
    """Generator over (task, [records per map]) for this rank's tasks.  Engine batches are bounded by the tile memory;
    a batch whose records overflow the engine's capacity is re-run with a larger one (MB200_ERR_CAPACITY contract).
    select=(pt, st): BH, the `o < pt` cut and the sparsity filter run on the device and the generator yields the selected
    candidates (ScaleSpaceEngine.candidates_batch) instead of every record."""
    from .engine import EngineError
    nmaps = 2 if differential else 1
    wc = min(dpx + 1, chunk - 1) - 3
    per_block = 2 * nmaps * chunk * wc * 8
    bmax = max(1, MAX_BATCH_TILE_BYTES // per_block)
    for first in range(0, len(tasks), bmax):
        batch = tasks[first:first + bmax]
        fraction = -1.0
        while True:
            eng.configure(chunk, dpx, nmaps * len(batch), record_fraction=fraction)
            if verbose:
                for t in batch:
                    print("Starting block ", t.block + 1, "/", (totals or {}).get(t.chrom, "?"), "...", sep="")
            eng.upload_coo_batch(0, *concat_coo([m for t in batch for m in t.maps]))
            if differential:
                eng.run_differential()
            else:
                eng.run()
            try:
                if select is not None:
                    eng.select_candidates(*select)
                    recs = eng.candidates_batch(pair=differential)
                else:
                    recs = eng.records_batch(pair=differential)
                break
            except EngineError as err:
                if err.code != -3 or fraction >= 1.0:
                    raise
                _, found = eng.batch_counts()
                fraction = min(1.0, max(0.25, 1.25 * float(found.max()) / float(chunk * wc)))
        if timings is not None:
            timings.append(eng.timing())
        for k, t in enumerate(batch):
            yield t, recs[nmaps * k:nmaps * (k + 1)]


def shard_and_call(preps, n_chrom, dpx, nmaps, eng, block_fn, rank=0, world=1, verbose=True, owners=None, timings=None,
                   width=4, select=None):
    """Runs every block of every chromosome once, somewhere, and returns {chromosome index: [call, ...]} on rank 0
    ({} elsewhere).  block_fn(task, records, chunk, start) -> list of calls [x, y, ..(width-2 more)] of that block in
    chromosome coordinates BEFORE the overlap de-duplication of process_block (mustache.py:945-960), applied here."""
    t0 = time.time()
    tasks, geom = build_tasks(preps, n_chrom, dpx, nmaps, rank, world, eng, owners)
    rows = []
    chunk = max(2 * dpx, 2000) if not geom else next(iter(geom.values()))[0]
    totals = {c: len(g[1]) for c, g in geom.items()}
    for task, recs in run_batches(eng, tasks, chunk, dpx, differential=(nmaps == 2), verbose=verbose, timings=timings,
                                  totals=totals, select=select):
        _, starts, ends = geom[task.chrom]
        ms = tiler.block_mask_size(task.block, starts, ends, dpx)
        for call in block_fn(task, recs, chunk, starts[task.block]):
            if tiler.keep_after_overlap(call, starts[task.block], ms):
                rows.append([task.chrom, task.block] + [float(a) for a in call])
        if verbose:
            print("Block", task.block + 1, "done.")
    arr = np.array(rows, dtype=np.float64).reshape(-1, width + 2)
    if world > 1:
        arr = sharding.gather_loops(arr, rank, world, collective_device(eng))
        if rank != 0:
            return {}
    order = np.lexsort((arr[:, 1], arr[:, 0])) if len(arr) else np.zeros(0, np.int64)   # stable: (chromosome, block)
    out = {}
    for r in arr[order]:
        out.setdefault(int(r[0]), []).append([int(r[2]), int(r[3])] + [float(a) for a in r[4:]])
    if timings is not None:
        timings.append({"host_s": time.time() - t0})
    return out
