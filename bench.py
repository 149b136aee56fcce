#!/usr/bin/env python3
"""Benchmark of the scale-space hot path (BASELINE.json metric: contact-bins/sec through the full scale-space).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5|chr|3s]

One JSON line on stdout (rank 0).  A "step" = one pass of the hot path (mask/fills, every Gaussian of every octave,
DoG, 3x3 maxima, extremum test, exponential-fit p-values, compact records) over one synthetic batch.

  --config 2 (default)  BASELINE configs[1]: synthetic 10k x 10k dense band (dpx 5000), 4 octaves x 12 sigma; with N > 1
                        every rank runs the same-shaped tile with its own seed (weak scaling).  The line also carries a
                        "config4" object: the 226-block workload of BASELINE configs[3] through the product's sharded
                        path at this N (strong scaling: compare the objects of the N = 1, 2, 4, 8 lines).
  --config 3 / 4 / 5    BASELINE configs[2] / [3] / [4] in full: the synthetic chromosomes of SURVEY 8(d) generated,
                        normalised and tiled exactly as the CLI does, the block pool spread over the N ranks by
                        mustache_b200.blockrun (owner ranks prepare, all_to_all exchange of block COO): strong scaling.
  --config chr / 3s     dense-band stand-ins of the 5 kb / 1 kb block shapes (24 x 2000^2, 6 x 4000^2).

`value`  : contact-bins/s with the tiles already resident in HBM, device time from CUDA events on the engine's stream,
           K steps bracketed by barrier + synchronize, max over ranks (the device-resident step has no collective: blocks
           are independent; `wall_ms_per_step` is the host clock around the same region).
`e2e`    : same metric through the public API with HOST buffers, the sequence the CLI runs: host -> device upload of every
           tile (dense pinned tile for config 2, block COO for the chromosome configs), all kernels, BH + `o < pt` + sparsity
           + enrichment filter on the device, device -> host read of the selected candidates, and for
           N > 1 the gather of the per-block result rows on the rank that writes the TSV.  Measured twice: with one engine
           handle (`single_engine_ms_per_step`: run, then post-processing and fetch, then the next run) and with two handles
           used alternately (mb200_run_after; the reported `e2e`): the post-processing and fetch of batch k overlap the run
           of batch k+1, timed from an empty pipeline to the last result on the host.
`--impl reference` times the UNMODIFIED reference (baseline/_ref, mustache.py:697-778 up to the intercepted
multipletests call) on all host cores over a bounded sample of the same workload; when baseline/_ref is absent, the
oracle port on the same scipy kernels.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_BIN_PER_OCTAVE = 272          # (12 reads + 11 writes + 11 reads) * 8 B
FP64_INSTR_PEAK = 1.849e13              # measured on this pool's B200 with tools/fp64_peak.cu (profiles/fp64_peak_r01.jsonl)

CONFIGS = {
    "2": dict(kind="dense", n=10000, dpx=5000, octaves=[1.6, 3.2, 6.4, 12.8], blocks=1,
              workload="synthetic 10k x 10k dense band (dpx 5000), 4 octaves x 12 sigma (BASELINE configs[1])"),
    "3s": dict(kind="dense", n=4000, dpx=2000, octaves=[1.6, 3.2], blocks=6,
               workload="6 dense-band blocks of 4000 x 4000 (dpx 2000), 2 octaves (1 kb block shape)"),
    "chr": dict(kind="dense", n=2000, dpx=400, octaves=[1.6, 3.2], blocks=24,
                workload="24 dense-band blocks of 2000 x 2000 (dpx 400), 2 octaves (5 kb block shape)"),
    "3": dict(kind="chrom", n=4000, dpx=2000, octaves=[1.6, 3.2], maps=1,
              workload="synthetic 50k-bin chromosome at 1 kb, Poisson(4/(d+1)) + 2000 loops, normalised, 24 blocks of "
                       "4000 x 4000 (dpx 2000), 2 octaves (BASELINE configs[2])"),
    "4": dict(kind="chrom", n=2000, dpx=400, octaves=[1.6, 3.2], maps=1,
              workload="8 synthetic chromosomes of 10k..80k bins at 5 kb, Poisson(18/(d+1)) + loops, normalised, 226 blocks "
                       "of 2000 x 2000 (dpx 400), 2 octaves (BASELINE configs[3])"),
    "5": dict(kind="chrom", n=2000, dpx=400, octaves=[1.6, 3.2], maps=2,
              workload="differential: two synthetic 20k-bin maps at 5 kb (map B = 0.6 thinning of A, 100 loops deleted, 100 "
                       "added), 13 block pairs x 3 filter stacks, 2 octaves (BASELINE configs[4])"),
}


def contact_bins(n, dpx):
    hi = min(dpx + 1, n - 1)
    return sum(n - k for k in range(4, hi + 1))


def fp64_instr_per_bin(octaves, dedupe=True):
    """FP64 instructions per contact-bin of the two separable passes: with the reference's arithmetic as it stands
    (scipy's folded taps, 3R+1 per output and pass), and as executed (the axis-0 pass shares the pair sums inside the
    groups mb_engine.cu:plan_kv cuts: R*(2n+1)+n per group of n steps with largest radius R; groups of at most 5, or at
    most 3 for chains that reach radius 24, which run the 64-register variant of the axis-0 kernel)."""
    from mustache_b200 import ladder
    prog = ladder.build_program(octaves, dedupe=dedupe)
    radii = sorted(s.radius for s in prog.steps)
    reference = sum(2 * (3 * r + 1) for r in radii)
    gmax, n = (3 if radii[-1] >= 24 else 5), len(radii)     # mb_kernels.cuh: KV_GSMALL / KV_GSMALL_RMIN / KV_GMAX
    best = [0] * (n + 1)
    for i in range(n - 1, -1, -1):
        best[i] = min(radii[i + c - 1] * (2 * c + 1) + c + best[i + c] for c in range(1, gmax + 1) if i + c <= n)
    executed = best[0] + sum(3 * r + 1 for r in radii)
    return reference, executed


def fp64_instr_per_bin_fma(octaves):
    """Executed FP64 instructions per contact-bin in the opt-in fast mode: a fused multiply-add per tap, i.e. R*(n+1)+n per
    axis-0 group of n steps and 2R+1 per axis-1 output."""
    from mustache_b200 import ladder
    radii = sorted(s.radius for s in ladder.build_program(octaves).steps)
    gmax, n = 5, len(radii)
    best = [0] * (n + 1)                                  # the grouping is the exact mode's (mb_engine.cu:plan_kv)
    take = [1] * (n + 1)
    for i in range(n - 1, -1, -1):
        best[i], take[i] = min((radii[i + c - 1] * (2 * c + 1) + c + best[i + c], c) for c in range(1, gmax + 1) if i + c <= n)
    kv, i = 0, 0
    while i < n:
        c = take[i]
        kv += radii[i + c - 1] * (c + 1) + c
        i += c
    return kv + sum(2 * r + 1 for r in radii)


def load_traffic():
    """DRAM bytes per launch of the kernels from the committed `ncu --set full` captures (profiles/traffic_r02.json, else
    round 1's), keyed by config name."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------
def make_host_tiles(cfg, rank):
    """Dense-band workloads in the reference-facing form: dense row-major N x N float64 tiles in pinned host memory."""
    from mustache_b200 import synth as gen
    from mustache_b200.engine import PinnedBuffer
    n, dpx = cfg["n"], cfg["dpx"]
    tiles, keep = [], []
    for b in range(cfg["blocks"]):
        band = gen.dense_band_tile(n, dpx, seed=1001 + 100 * rank + 7 * b, blob_seed=1002 + 100 * rank + 7 * b,
                                   nblobs=200 if n >= 4000 else 40)
        buf = PinnedBuffer((n, n))
        keep.append(buf)
        dense = buf.array
        dense[:] = 0.0
        w = band.shape[1]
        safe = max(0, min(n, n - 4 - w + 1))
        if safe > 0:
            view = np.lib.stride_tricks.as_strided(dense.ravel()[4:], shape=(safe, w), strides=((n + 1) * 8, 8))
            view[:] = band[:safe]
        for i in range(safe, n):
            ww = min(w, n - i - 4)
            if ww > 0:
                dense[i, i + 4:i + 4 + ww] = band[i, :ww]
        tiles.append(dense)
    return tiles, keep


def chromosome_specs(name):
    """[(chromosome name, generator kwargs, res)] of a chromosome config, as tests/golden/make_golden.py fed the reference."""
    from mustache_b200 import synth as gen
    if name == "3":
        return [("chrS", dict(gen.CONFIG3))]
    if name == "4":
        return [(k, dict(v)) for k, v in gen.CONFIG4.items()]
    if name == "5":
        return [("chrD", dict(gen.CONFIG5))]
    raise KeyError(name)


def prepare_owned(name, rank, world):
    """What the CLI's owner ranks do for their chromosomes (read + normalise), minus the text round trip: generate the
    raw counts, normalise with the host normaliser the CLI uses.  Returns (preps, n_chrom, owners)."""
    from mustache_b200 import sharding, synth as gen
    from mustache_b200.normalize import normalize_sparse
    specs = chromosome_specs(name)
    owners = sharding.chromosome_owners(len(specs), world, sizes=[s["n"] for _, s in specs])
    preps = {}
    for c, (_, spec) in enumerate(specs):
        if owners[c] != rank:
            continue
        res = spec.pop("res")
        if name == "5":
            maps = [(x, y, v.astype(np.float64)) for x, y, v in gen.config5_maps(**spec)]
        else:
            x, y, v = gen.synthetic_chromosome(**spec)
            maps = [(x, y, v.astype(np.float64))]
        for x, y, v in maps:
            normalize_sparse(x, y, v, res, spec["dpx"])
        preps[c] = dict(maps=maps, n=spec["n"])
    return preps, len(specs), owners


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    n, dpx, octaves, seed, use_reference = args
    from mustache_b200 import synth as gen
    c = gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=seed, blob_seed=seed + 1, nblobs=10), n)
    if use_reference:
        from baseline import reference_arm
        return reference_arm.time_scale_space(c, dpx, octaves)
    from oracle import scalespace as osc
    t0 = time.perf_counter()
    res = osc.scale_space(c, dpx, octaves, use_scipy=True)
    return time.perf_counter() - t0, int((res["p"] != 2).sum())


def cpu_sample_geometry(cfg, target_core_seconds=12.0):
    """Sub-tile with the same band-to-tile ratio as the workload's blocks, sized for ~target seconds per core."""
    per_bin = 1.0 / (37.4e3 if len(cfg["octaves"]) >= 4 else 88.6e3)        # SURVEY.md section 6, per core
    if contact_bins(cfg["n"], cfg["dpx"]) * per_bin <= 1.5 * target_core_seconds:
        return cfg["n"], cfg["dpx"]                                          # a whole block of the workload fits the budget
    ratio = cfg["dpx"] / cfg["n"]
    n = 400
    while n < cfg["n"]:
        if contact_bins(n + 100, int((n + 100) * ratio)) * per_bin > target_core_seconds:
            break
        n += 100
    return n, max(8, int(n * ratio))


def run_cpu_arm(cfg, steps, warmup, cores=None):
    """The reference's scale-space loop over Pool(cores), one tile per process and step.  Returns value, ms, cores,
    kind, sample description."""
    import multiprocessing as mp
    from baseline import reference_arm
    use_ref = reference_arm.available()
    cores = cores or os.cpu_count() or 1
    n, dpx = cpu_sample_geometry(cfg)
    bins = contact_bins(n, dpx) * cores
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(n, dpx, cfg["octaves"], 5000 + 31 * it + k, use_ref) for k in range(cores)])
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    dt = float(np.mean(times))
    what = ("unmodified reference mustache() from entry to its multipletests call (baseline/_ref, mustache.py:697-778)"
            if use_ref else "scipy path of the oracle port (baseline/_ref absent)")
    sample = "%d tiles of %d x %d (dpx %d, %d octaves) per step, one per process over Pool(%d); %s" % (
        cores, n, n, dpx, len(cfg["octaves"]), cores, what)
    return bins / dt, dt * 1e3, cores, ("reference" if use_ref else "port"), sample


# ------------------------------------------------------------------------------------------------------------
# measurement
# ------------------------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, rank, local_rank, world):
        import torch
        self.torch, self.rank, self.local_rank, self.world = torch, rank, local_rank, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def measure(h, eng, cfg, name, steps, warmup, device_only=False, eng2=None):
    """Device-resident and end-to-end timing of one workload on this rank's engine.  Returns a dict of raw figures."""
    from mustache_b200 import blockrun, sharding
    rank, world = h.rank, h.world
    eng.set_octaves(cfg["octaves"], differential=(cfg.get("maps", 1) == 2))
    n, dpx = cfg["n"], cfg["dpx"]
    nmaps = cfg.get("maps", 1)
    if cfg["kind"] == "dense":
        tiles, keep = make_host_tiles(cfg, rank)
        nblk = cfg["blocks"]

        def upload():
            for b, t in enumerate(tiles):
                eng.upload_dense(b, t)
        wc = min(dpx + 1, n - 1) - 3
        h2d = nblk * n * wc * 8
        stacks = 1
    else:
        preps, n_chrom, owners = prepare_owned(name, rank, world)
        tasks, geom = blockrun.build_tasks(preps, n_chrom, dpx, nmaps, rank, world, eng, owners)
        nblk = nmaps * len(tasks)
        # block COO of the whole batch, concatenated, in page-locked host memory (what a caller that wants full-speed
        # uploads hands to mb200_upload_coo_batch)
        from mustache_b200.engine import PinnedBuffer
        offsets, rows, cols, vals = blockrun.concat_coo([m for t in tasks for m in t.maps])
        keep = [PinnedBuffer((max(len(vals), 1),), np.int32), PinnedBuffer((max(len(vals), 1),), np.int32),
                PinnedBuffer((max(len(vals), 1),), np.float64)]
        flat = [buf.array[:len(vals)] for buf in keep]
        for dst, src in zip(flat, (rows, cols, vals)):
            dst[:] = src

        def upload():
            if nblk:
                eng.upload_coo_batch(0, offsets, *flat)
        h2d = 16 * len(vals)
        stacks = 3 if nmaps == 2 else 1                  # SURVEY 8(d) counts the three filter stacks of a block pair
    bins_rank = contact_bins(n, dpx) * (nblk // nmaps) * stacks
    run = eng.run_differential if nmaps == 2 else eng.run
    eng.configure(n, dpx, max(nblk, nmaps))
    upload()
    eng.sync()

    for _ in range(warmup):
        run()
    h.barrier()
    dev_ms, phases = 0.0, {"prep_ms": 0.0, "kv_ms": 0.0, "kh_ms": 0.0, "ks_ms": 0.0, "fin_ms": 0.0}
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
        eng.sync()
        tm = eng.timing()
        dev_ms += tm["total_ms"]
        for k in phases:
            phases[k] += tm[k]
    h.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = eng.launches() * steps
    step_ms = h.max_over_ranks(dev_ms / steps)        # CUDA events on the engine's stream, every N; max over ranks
    wall_step_ms = h.max_over_ranks(wall_ms / steps)
    bins_all = h.sum_over_ranks(bins_rank)
    out = dict(step_ms=step_ms, bins_rank=bins_rank, bins_all=bins_all, value=bins_all / (step_ms * 1e-3), launches=launches,
               phases={k: v / steps for k, v in phases.items()}, dev_ms=dev_ms / steps, blocks_rank=nblk // nmaps,
               wall_step_ms=wall_step_ms)
    if device_only:
        return out

    # ---- end to end through the public API with host buffers ----
    # Every step: host tiles -> device, all kernels, BH + selection + filters on the device, ONE fetch of the selected
    # candidates, and for N > 1 the gather of the per-block result rows on rank 0.  The uploads of the next batch are issued
    # right after a run (the engine double-buffers tiles on a second stream), which is how a caller with more than one batch
    # uses the API.
    dev = blockrun.collective_device(eng)
    sel = (0.05 if nmaps == 2 else 0.1, 0.88)

    def fetch(e):
        # the CLI's path: BH + o < pt + sparsity + enrichment filter on the device, only the selected candidates (with the
        # neighbourhoods the clustering and, for config 5, the differential selection read) come back
        e.select_candidates(*sel)
        recs = e.candidates_batch(pair=(nmaps == 2))
        rows = np.array([[rank, b, r["n_found"], r["nz_count"]] for b, r in enumerate(recs)], dtype=np.float64).reshape(-1, 4)
        got = sharding.gather_loops(rows, rank, world, dev) if world > 1 else rows
        return recs, got

    def e2e_step():
        run()
        upload()
        return fetch(eng)

    upload()
    for _ in range(max(1, warmup // 2)):
        recs, got = e2e_step()
    h.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        recs, got = e2e_step()
    h.barrier()
    e2e_single_ms = h.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    e2e_ms, e2e_mode = e2e_single_ms, "one engine: run, then post-processing and fetch, then the next run"

    # Two engine handles on the GPU used alternately (mb200_run_after): the run of batch k+1 follows the run of batch k back
    # to back while BH / selection / filters and the candidate fetch of batch k overlap it.  The timed region starts with an
    # empty pipeline and ends with every result on the host: `steps` uploads, runs and fetches.
    ready = 0
    if eng2 is not None:
        try:                                               # the second engine's scratch is allocated here
            eng2.set_octaves(cfg["octaves"], differential=(nmaps == 2))
            eng2.configure(n, dpx, max(nblk, nmaps))
            ready = 1
        except Exception as err:
            e2e_mode += " (second engine unavailable: %s)" % str(err)[:120]
    # every rank takes the same path (the fetches below contain a collective): all of them pipeline, or none does
    if -h.max_over_ranks(-ready) > 0:
        try:
            engines = [eng, eng2]

            def upload_to(e):
                if cfg["kind"] == "dense":
                    for b, t in enumerate(tiles):
                        e.upload_dense(b, t)
                elif nblk:
                    e.upload_coo_batch(0, offsets, *flat)

            trace = [] if os.environ.get("BENCH_TRACE") else None

            def pipelined(count):
                out = None
                for i in range(count + 1):
                    e, o = engines[i % 2], engines[(i + 1) % 2]
                    ta = time.perf_counter()
                    if i < count:
                        if i > 0:
                            e.run_after(o)
                        (e.run_differential if nmaps == 2 else e.run)()
                        tb = time.perf_counter()
                        upload_to(e)                       # tiles of the batch this engine runs next
                    else:
                        tb = ta
                    tc = time.perf_counter()
                    if i > 0:
                        out = fetch(o)                     # results of step i - 1, next to the run of step i
                    if trace is not None:
                        td = time.perf_counter()
                        trace.append((round((tb - ta) * 1e3, 3), round((tc - tb) * 1e3, 3), round((td - tc) * 1e3, 3),
                                      round(o.timing()["total_ms"], 3) if i > 0 else None, round(o.post_ms(), 3) if i > 0 else None))
                return out

            for e in engines:
                upload_to(e)
            pipelined(max(2, warmup // 2))
            for e in engines:
                e.sync()
            h.barrier()
            t0 = time.perf_counter()
            recs, got = pipelined(steps)
            h.barrier()
            e2e_ms = h.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
            e2e_mode = ("two engine handles used alternately (mb200_run_after): post-processing and fetch of batch k overlap "
                        "the run of batch k+1; timed from an empty pipeline to the last result on the host")
            eng2.sync()
            if trace is not None:
                sys.stderr.write("pipeline trace (run enqueue, upload enqueue, fetch, device run ms, post ms): %s\n" % trace[-(steps + 1):])
        except Exception as err:
            if world > 1:                                  # the other ranks are inside the collectives of the pipeline
                raise
            e2e_mode += " (two-engine pipeline failed: %s)" % str(err)[:120]
            eng.configure(n, dpx, max(nblk, nmaps))
            upload()
            run()
            recs, got = fetch(eng)
    eng.sync()
    n_found = int(sum(r["n_found"] for r in recs))
    total_found = h.sum_over_ranks(n_found)
    if rank == 0:
        assert int(got[:, 2].sum()) == int(total_found), "rank 0 did not receive every block's result row"
    n_cand = int(sum(len(r["rows"]) for r in recs))
    # block, row, col, flags; q, sigma, cval; o9, so9 (+ pair9, vself9, vother9 for the differential path); counters
    d2h = n_cand * (4 * 4 + 3 * 8 + (45 if nmaps == 2 else 18) * 8) + nblk * 16 + 8
    post_ms = eng.post_ms()
    out.update(e2e_ms=e2e_ms, e2e_value=bins_all / (e2e_ms * 1e-3), e2e_single_ms=e2e_single_ms, e2e_mode=e2e_mode,
               h2d=int(h.sum_over_ranks(h2d)),
               d2h=int(h.sum_over_ranks(d2h)), n_found=int(total_found), post_ms=post_ms,
               n_candidates=None if n_cand is None else int(h.sum_over_ranks(n_cand)))
    return out


def bench_normaliser(eng, steps, warmup):
    """normalize_sparse (mustache.py:622-686) on the device, the producer of the tile values.  Inputs:
    the BASELINE configs[2] chromosome (1 kb, 2 000-bin windows) and a 5 kb chromosome of configs[3]; bias-corrected counts.
    Timed through the C ABI with host arrays (H2D of x, y, v and D2H of v inside); the numpy normaliser the CLI would
    otherwise run (the reference's own code path, one core) beside it on the same input."""
    from mustache_b200 import synth as gen
    from mustache_b200.normalize import normalize_sparse, normalize_sparse_device
    out = {}
    for name, spec in (("config3_1kb", dict(gen.CONFIG3)), ("config4_s8_5kb", dict(gen.CONFIG4["s8"]))):
        res = spec.pop("res")
        x, y, c = gen.synthetic_chromosome(**spec)
        bias = np.random.default_rng(5).uniform(0.6, 1.6, size=spec["n"])
        v0 = c / bias[x] / bias[y]
        x32, y32 = x.astype(np.int32), y.astype(np.int32)
        for _ in range(warmup):
            normalize_sparse_device(eng, x32, y32, v0.copy(), res, spec["dpx"])
        ts = []
        for _ in range(steps):
            v = v0.copy()
            t0 = time.perf_counter()
            normalize_sparse_device(eng, x32, y32, v, res, spec["dpx"])
            ts.append(time.perf_counter() - t0)
        dev = v
        t0 = time.perf_counter()
        ref = v0.copy()
        normalize_sparse(x, y, ref, res, spec["dpx"])
        cpu_s = time.perf_counter() - t0
        ms = float(np.median(ts)) * 1e3
        out[name] = {"contacts": int(len(v0)), "bins": spec["n"], "diagonals": spec["dpx"] + 2, "window_bins": 2000000 // res,
                     "ms": ms, "contacts_per_s": len(v0) / (ms * 1e-3), "bytes_per_s": 24 * len(v0) / (ms * 1e-3),
                     "launches": eng.launches(), "numpy_one_core_s": cpu_s, "speedup_vs_numpy": cpu_s / (ms * 1e-3),
                     "max_abs_diff_vs_numpy": float(np.abs(dev - ref).max())}
    return out


def roofline(cfg, name, m, peak, peaks_found):
    n_oct = len(cfg["octaves"])
    bytes_per_bin = BYTES_PER_BIN_PER_OCTAVE * n_oct
    ph = m["phases"]
    fused = ph["ks_ms"] < 0.01                          # khs_kernel: axis-1 pass, DoG and scoring in one kernel (2-octave chains)
    khn = "khs_kernel" if fused else "kh_kernel"
    kern = {"kv_kernel": ph["kv_ms"], khn: ph["kh_ms"], "ks_kernel": ph["ks_ms"]}
    dom = max(kern, key=kern.get)
    achieved = m["bins_rank"] * bytes_per_bin / (m["dev_ms"] * 1e-3) / 1e9
    instr_ref, instr_exec = fp64_instr_per_bin(cfg["octaves"])
    hot_ms = ph["kv_ms"] + ph["kh_ms"]
    share = {"kv_kernel": 12 * 8 * n_oct, "kh_kernel": 11 * 8 * n_oct, "ks_kernel": 11 * 8 * n_oct, "khs_kernel": 22 * 8 * n_oct}
    tr = (load_traffic() or {}).get(name, {})
    per_kernel = {}
    for kname, key in (("kv_kernel", "kv_ms"), (khn, "kh_ms"), ("ks_kernel", "ks_ms")):
        if ph[key] < 0.01:
            continue
        ach = m["bins_rank"] * share[kname] / (ph[key] * 1e-3) / 1e9
        per_kernel[kname] = {"ms": ph[key], "algorithmic_bytes_per_bin": share[kname], "achieved": ach, "frac": ach / peak,
                             "traffic": tr.get(kname)}
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": tr.get(dom), "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks_found else "fallback",
            "algorithmic_bytes_per_bin": bytes_per_bin, "bins_per_launch": m["bins_rank"],
            "scope": "whole step on rank 0 (prep + axis-0 + axis-1/DoG + scoring + statistics), device time from CUDA events",
            "dominant_kernel": dom, "dominant_kernel_share": kern[dom] / max(m["dev_ms"], 1e-9),
            "traffic_note": ("dram__bytes_read+write of the dominant kernel per launch, ncu --set full capture of this command "
                             "(profiles/)") if tr else "no ncu capture committed for this config",
            "kernels": per_kernel,
            "fp64": {"instr_per_bin_reference": instr_ref, "instr_per_bin_executed": instr_exec,
                     "achieved_instr_per_s": instr_exec * m["bins_rank"] / max(hot_ms * 1e-3, 1e-12),
                     "peak_instr_per_s": FP64_INSTR_PEAK,
                     "frac": instr_exec * m["bins_rank"] / max(hot_ms * 1e-3, 1e-12) / FP64_INSTR_PEAK,
                     "note": "axis-0 + axis-1 kernels; an FP64 instruction holds the SM sub-partition's dispatch for 2 cycles, "
                             "every other instruction costs ~0.75 more (tools/fp64_peak.cu)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=list(CONFIGS) + ["norm"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the config-4 strong-scaling object of multi-GPU lines")
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--fusion", type=int, default=0, help="mb200_set_fusion: 1 = axis-1 + scoring fused, 2 = axis-0 + axis-1 fused")
    ap.add_argument("--overlap", action="store_true", help="two half-batches on two streams (mb200_set_overlap)")
    ap.add_argument("--no-fast", action="store_true", help="skip the extra device-resident measurement in the opt-in FMA mode")
    ap.add_argument("--single-engine", action="store_true", help="e2e with one engine handle only (no pipelining of the post-processing)")
    ap.add_argument("--device-only", action="store_true", help="profiling aid: only the device-resident steps (no e2e, no CPU leg)")
    args = ap.parse_args()
    if args.config == "norm":
        from mustache_b200.engine import ScaleSpaceEngine
        print(json.dumps({"metric": "normalize_sparse_contacts_per_sec", "unit": "contacts/s", "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "dtype": "f64", "data": "synthetic",
                          "inputs": bench_normaliser(ScaleSpaceEngine(int(os.environ.get("LOCAL_RANK", "0"))), args.steps, args.warmup)}))
        return
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = cfg["kind"] == "chrom"
    config = {"workload": cfg["workload"], "n": cfg["n"], "dpx": cfg["dpx"], "octaves": cfg["octaves"],
              "l2": "inputs larger than L2 (band tiles + axis-0 scratch >> 126 MB)",
              "parallelism": ("block pool of all chromosomes spread evenly over the ranks (owner ranks prepare, all_to_all of "
                              "block COO), per-rank post-processing, gather of result rows to rank 0" if strong else
                              "one tile set per GPU, gather of result rows to rank 0") if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        val, ms, cores, kind, sample = run_cpu_arm(cfg, steps, warm)
        config["workload"] = cfg["workload"] + " -- CPU arm sampled as: " + sample
        print(json.dumps({"impl": "reference", "metric": "contact_bins_per_sec", "value": val, "unit": "contact-bins/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": val, "unit": "contact-bins/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": "contact-bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mustache_b200.engine import ScaleSpaceEngine
    h = Harness(rank, local_rank, world)
    eng = ScaleSpaceEngine(local_rank)
    if args.overlap:
        eng.set_overlap(True)
        config["overlap"] = "two half-batches on two streams"
    if args.fusion:
        eng.set_fusion(args.fusion)
        config["fusion"] = {1: "khs_kernel (axis-1 + DoG + scoring in one kernel)", 2: "kvh_kernel (axis-0 + axis-1 + DoG in one kernel)"}[args.fusion]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # 200 ms period: started before the warm-up so that short timed regions are covered
    eng2 = None if (args.device_only or args.single_engine) else ScaleSpaceEngine(local_rank)   # second handle for the e2e pipeline
    m = measure(h, eng, cfg, args.config, args.steps, args.warmup, device_only=args.device_only, eng2=eng2)
    clocks = sampler.stop() if rank == 0 else None
    if args.device_only:
        if rank == 0:
            print(json.dumps({"device_only": True, "ms_per_step": m["step_ms"], "value": m["value"],
                              "phases_ms_per_step": m["phases"]}))
        return
    fastm = None
    if not args.no_fast and world == 1:
        eng.set_arithmetic(True)
        fastm = measure(h, eng, cfg, args.config, args.steps, args.warmup, device_only=True)
        eng.set_arithmetic(False)
    c4 = None
    if not strong and not args.no_config4:
        c4m = measure(h, eng, CONFIGS["4"], "4", max(3, args.steps), 3, eng2=eng2)
        c4 = {"workload": CONFIGS["4"]["workload"], "scaling": "strong", "n_gpus": world, "value": c4m["value"],
              "unit": "contact-bins/s", "ms_per_step": c4m["step_ms"], "blocks_on_rank0": c4m["blocks_rank"],
              "e2e": {"value": c4m["e2e_value"], "ms_per_step": c4m["e2e_ms"], "h2d_bytes_per_step": c4m["h2d"],
                      "d2h_bytes_per_step": c4m["d2h"], "single_engine_ms_per_step": c4m["e2e_single_ms"]},
              "records_per_step": c4m["n_found"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    config["blocks_on_rank0"] = m["blocks_rank"]
    out = {"metric": "contact_bins_per_sec", "value": m["value"], "unit": "contact-bins/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": m["step_ms"], "higher_is_better": True,
           "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "clocks": clocks,
           "e2e": {"value": m["e2e_value"], "unit": "contact-bins/s", "ms_per_step": m["e2e_ms"],
                   "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"], "mode": m["e2e_mode"],
                   "single_engine_ms_per_step": m["e2e_single_ms"]},
           "gpu_launches": m["launches"], "records_per_step": m["n_found"], "candidates_per_step": m["n_candidates"],
           "post_ms_per_step": m["post_ms"], "phases_ms_per_step": m["phases"],
           "wall_ms_per_step": m["wall_step_ms"],
           "roofline": roofline(cfg, args.config, m, peak, bool(peaks))}
    if fastm is not None:
        fi = fp64_instr_per_bin_fma(cfg["octaves"])
        hot = (fastm["phases"]["kv_ms"] + fastm["phases"]["kh_ms"]) * 1e-3
        out["fast_mode"] = {"note": "opt-in mb200_set_arithmetic(1) / MUSTACHE_FAST=1: one FMA per tap instead of scipy's multiply-then-add; "
                                    "not the reference's arithmetic, gated by tests/test_gpu_fast_mode.py on BASELINE's tolerance; "
                                    "every other figure of this line is the exact mode",
                            "ms_per_step": fastm["step_ms"], "value": fastm["value"], "phases_ms_per_step": fastm["phases"],
                            "hbm_frac": fastm["bins_rank"] * BYTES_PER_BIN_PER_OCTAVE * len(cfg["octaves"]) / (fastm["dev_ms"] * 1e-3) / 1e9 / peak,
                            "fp64": {"instr_per_bin_executed": fi, "achieved_instr_per_s": fi * fastm["bins_rank"] / hot,
                                     "frac": fi * fastm["bins_rank"] / hot / FP64_INSTR_PEAK}}
    if c4 is not None:
        out["config4"] = c4
    if not args.no_cpu_baseline and world == 1:
        val, ms, cores, kind, sample = run_cpu_arm(cfg, args.cpu_steps, 0)
        out["cpu_baseline"] = {"value": val, "unit": "contact-bins/s", "cores": cores, "kind": kind, "sample": sample,
                               "ms_per_sample": ms}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
