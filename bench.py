#!/usr/bin/env python3
"""Benchmark of the scale-space hot path (BASELINE.json metric: contact-bins/sec through the full scale-space).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3s|chr]

One JSON line on stdout (rank 0).  A "step" = one pass of the hot path (mask/fills, every Gaussian of every octave,
DoG, 3x3 maxima, extremum test, exponential-fit p-values, compact records) over one synthetic batch:
  N=1 default workload = BASELINE.json configs[1]: synthetic 10k x 10k dense band (dpx 5000), 4 octaves x 12 sigma.
  N>1: every rank runs the same-shaped tile with its own seed (weak scaling); the only collective is the NCCL
  gather of the candidate records to the rank that runs BH-FDR, inside the timed step.
`value`  : contact-bins/s with the tile already resident in HBM, device time from CUDA events on the engine's stream.
`e2e`    : same metric through the public API with the HOST tile: pinned host -> device copy of the band, all kernels,
           device -> host read of the records, per step, wall clock around synchronised calls.
`--impl reference` times the reference's CPU algorithm (oracle port on scipy, i.e. the same scipy.ndimage C kernels
mustache.py calls) on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import math
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
    # name: (n, dpx, octaves, blocks per rank, description)
    "2": dict(n=10000, dpx=5000, octaves=[1.6, 3.2, 6.4, 12.8], blocks=1,
              workload="synthetic 10k x 10k dense band (dpx 5000), 4 octaves x 12 sigma (BASELINE configs[1])"),
    "3s": dict(n=4000, dpx=2000, octaves=[1.6, 3.2], blocks=6,
               workload="6 blocks of a synthetic 1kb-style band (N 4000, dpx 2000), 2 octaves (slice of configs[2])"),
    "chr": dict(n=2000, dpx=400, octaves=[1.6, 3.2], blocks=24,
                workload="24 dense-band blocks of 2000 x 2000 (dpx 400), 2 octaves (5 kb chromosome shape, configs[3])"),
}


def contact_bins(n, dpx):
    hi = min(dpx + 1, n - 1)
    return sum(n - k for k in range(4, hi + 1))


def fp64_instr_per_bin(octaves, dedupe=True):
    """FP64 instructions per contact-bin of the two separable passes: with the reference's arithmetic as it stands
    (scipy's folded taps, 3R+1 per output and pass), and as executed (the axis-0 pass shares the pair sums inside the
    groups mb_engine.cu:plan_kv cuts: R*(2n+1)+n per group of n steps with largest radius R)."""
    from mustache_b200 import ladder
    prog = ladder.build_program(octaves, dedupe=dedupe)
    radii = sorted(s.radius for s in prog.steps)
    reference = sum(2 * (3 * r + 1) for r in radii)
    gmax, n = 5, len(radii)
    best = [0] * (n + 1)
    for i in range(n - 1, -1, -1):
        best[i] = min(radii[i + c - 1] * (2 * c + 1) + c + best[i + c] for c in range(1, gmax + 1) if i + c <= n)
    executed = best[0] + sum(3 * r + 1 for r in radii)
    return reference, executed


def load_traffic():
    """DRAM bytes per launch of the three kernels from the committed `ncu --set full` capture of this workload
    (profiles/traffic_r01.json, written by tools/ncu_summary.py --traffic); None when the capture is for another config."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
    except Exception:
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


def make_host_tiles(cfg, rank, pinned=True):
    """Synthetic tiles in the reference-facing form: dense row-major N x N float64 (pinned host memory)."""
    from mustache_b200 import synth as gen
    from mustache_b200.engine import PinnedBuffer
    n, dpx = cfg["n"], cfg["dpx"]
    tiles, keep = [], []
    for b in range(cfg["blocks"]):
        band = gen.dense_band_tile(n, dpx, seed=1001 + 100 * rank + 7 * b, blob_seed=1002 + 100 * rank + 7 * b,
                                   nblobs=200 if n >= 4000 else 40)
        if pinned:
            buf = PinnedBuffer((n, n))
            keep.append(buf)
            dense = buf.array
            dense[:] = 0.0
        else:
            dense = np.zeros((n, n))
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


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (scipy path), all host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    n, dpx, octaves, seed = args
    from mustache_b200 import synth as gen
    from oracle import scalespace as osc
    c = gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=seed, blob_seed=seed + 1, nblobs=10), n)
    t0 = time.perf_counter()
    res = osc.scale_space(c, dpx, octaves, use_scipy=True)
    return time.perf_counter() - t0, int((res["p"] != 2).sum())


def cpu_sample_geometry(cfg, target_core_seconds=12.0):
    """Sub-tile with the same band-to-tile ratio as the workload, sized for ~target seconds per core."""
    per_bin = 1.0 / (37.4e3 if len(cfg["octaves"]) >= 4 else 88.6e3)        # SURVEY.md section 6, per core
    ratio = cfg["dpx"] / cfg["n"]
    n = 400
    while n < cfg["n"]:
        if contact_bins(n + 100, int((n + 100) * ratio)) * per_bin > target_core_seconds:
            break
        n += 100
    return n, max(8, int(n * ratio))


def run_cpu_port(cfg, steps, warmup, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n, dpx = cpu_sample_geometry(cfg)
    bins = contact_bins(n, dpx) * cores
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(n, dpx, cfg["octaves"], 5000 + 31 * it + k) for k in range(cores)])
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    dt = float(np.mean(times))
    sample = "%d tiles of %d x %d (dpx %d, %d octaves), one per process, scipy path of the oracle port" % (
        cores, n, n, dpx, len(cfg["octaves"]))
    return bins / dt, dt * 1e3, cores, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--device-only", action="store_true", help="profiling aid: only the device-resident steps (no e2e, no CPU leg)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_oct = len(cfg["octaves"])
    bytes_per_bin = BYTES_PER_BIN_PER_OCTAVE * n_oct
    config = {"workload": cfg["workload"], "n": cfg["n"], "dpx": cfg["dpx"], "octaves": cfg["octaves"],
              "blocks_per_gpu": cfg["blocks"], "l2": "inputs larger than L2 (band tile + axis-0 scratch >> 126 MB)",
              "parallelism": "blocks sharded one set per GPU, NCCL gather of the records to rank 0 only" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        val, ms, cores, sample = run_cpu_port(cfg, steps, warm)
        print(json.dumps({"impl": "reference", "metric": "contact_bins_per_sec", "value": val, "unit": "contact-bins/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "contact-bins/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "contact-bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mustache_b200.engine import ScaleSpaceEngine
    from mustache_b200 import gather

    eng = ScaleSpaceEngine(local_rank)
    eng.set_octaves(cfg["octaves"])
    tiles, keep = make_host_tiles(cfg, rank)
    n, dpx, B = cfg["n"], cfg["dpx"], cfg["blocks"]
    bins_rank = contact_bins(n, dpx) * B
    eng.configure(n, dpx, B)
    for b, t in enumerate(tiles):
        eng.upload_dense(b, t)
    eng.sync()

    def device_step():
        eng.run()
        if world > 1:                        # the path's only collective: candidate records, device to device over NCCL
            for b in range(B):
                gather.gather_device_to_root(eng.records_device(b), world, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # 200 ms period: started before the warm-up so that short timed regions are covered
    for _ in range(args.warmup):
        device_step()
    barrier()
    dev_ms, phases = 0.0, {"prep_ms": 0.0, "kv_ms": 0.0, "kh_ms": 0.0, "ks_ms": 0.0, "fin_ms": 0.0}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        device_step()
        eng.sync()
        tm = eng.timing()
        dev_ms += tm["total_ms"]
        for k in phases:
            phases[k] += tm[k]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches() * args.steps
    step_ms = (wall_ms if world > 1 else dev_ms) / args.steps
    if world > 1:
        tt = torch.tensor([step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item())
    value = bins_rank * world / (step_ms * 1e-3)

    if args.device_only:
        if rank == 0:
            print(json.dumps({"device_only": True, "ms_per_step": step_ms, "value": value,
                              "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()}}))
        return

    # ---- end to end through the public API with host buffers ----
    # Every step: pinned host tiles -> device (band only), all kernels, records device -> host.  The uploads of step k+1
    # are issued right after mb200_run of step k (the engine double-buffers tiles on a second stream), which is how a
    # caller with more than one batch uses the API; K steps = K uploads + K runs + K record fetches inside the region.
    def e2e_upload():
        for b, t in enumerate(tiles):
            eng.upload_dense(b, t)

    def e2e_step():
        eng.run()
        e2e_upload()
        if world > 1:
            for b in range(B):
                gather.gather_device_to_root(eng.records_device(b), world, rank)
        recs = [eng.records(b, sort=False, pinned=(B == 1)) for b in range(B)]
        return recs

    e2e_upload()
    for _ in range(max(1, args.warmup // 2)):
        recs = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        recs = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    eng.sync()
    if world > 1:
        tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    wc = min(dpx + 1, n - 1) - 3
    h2d = B * n * wc * 8
    d2h = int(sum(r["n_found"] for r in recs)) * 36 + B * 20      # rows, cols, score id (int32), v, p, sigma (float64) + counters
    n_found = int(sum(r["n_found"] for r in recs))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hot_ms = (phases["kv_ms"] + phases["kh_ms"]) / args.steps
    kern = {"kv_kernel": phases["kv_ms"], "kh_kernel": phases["kh_ms"], "ks_kernel": phases["ks_ms"]}
    dom = max(kern, key=kern.get)
    achieved = bins_rank * bytes_per_bin / (dev_ms / args.steps * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    instr_ref, instr_exec = fp64_instr_per_bin(cfg["octaves"])
    instr = instr_exec * bins_rank
    steps_ms = {k: v / args.steps for k, v in phases.items()}
    # SURVEY 8(d) accounting split by the kernel that moves the bytes: 12 input reads (axis-0 pass), 11 DoG writes
    # (axis-1 pass), 11 DoG reads (scoring) per octave
    share = {"kv_kernel": 12 * 8 * n_oct, "kh_kernel": 11 * 8 * n_oct, "ks_kernel": 11 * 8 * n_oct}
    traffic = load_traffic()
    tr = (traffic or {}).get(args.config, {})
    per_kernel = {}
    for kname, ph in (("kv_kernel", "kv_ms"), ("kh_kernel", "kh_ms"), ("ks_kernel", "ks_ms")):
        ach = bins_rank * share[kname] / (steps_ms[ph] * 1e-3) / 1e9
        per_kernel[kname] = {"ms": steps_ms[ph], "algorithmic_bytes_per_bin": share[kname], "achieved": ach,
                             "frac": ach / peak, "traffic": tr.get(kname)}
    out = {"metric": "contact_bins_per_sec", "value": value, "unit": "contact-bins/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config,
           "clocks": clocks,
           "e2e": {"value": bins_rank * world / (e2e_ms * 1e-3), "unit": "contact-bins/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": launches,
           "records_per_step": n_found,
           "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": tr.get(dom), "peak_source": "measured" if peaks else "fallback",
                        "algorithmic_bytes_per_bin": bytes_per_bin, "bins_per_launch": bins_rank,
                        "scope": "whole step (prep + kv_kernel + kh_kernel + ks_kernel + statistics), device time from CUDA events",
                        "dominant_kernel": dom, "dominant_kernel_share": kern[dom] / max(dev_ms, 1e-9),
                        "traffic_note": "dram__bytes_read+write of the dominant kernel per launch, ncu --set full capture "
                                        "of this command (profiles/)" if tr else "no ncu capture committed for this config",
                        "kernels": per_kernel,
                        "fp64": {"instr_per_bin_reference": instr_ref, "instr_per_bin_executed": instr_exec,
                                 "achieved_instr_per_s": instr / (hot_ms * 1e-3), "peak_instr_per_s": FP64_INSTR_PEAK,
                                 "frac": instr / (hot_ms * 1e-3) / FP64_INSTR_PEAK,
                                 "note": "kv_kernel + kh_kernel; an FP64 instruction holds the SM sub-partition's dispatch "
                                         "for 2 cycles, every other instruction costs ~0.75 more (tools/fp64_peak.cu)"}}}
    if not args.no_cpu_baseline and world == 1:
        val, ms, cores, sample = run_cpu_port(cfg, args.cpu_steps, 0)
        out["cpu_baseline"] = {"value": val, "unit": "contact-bins/s", "cores": cores, "kind": "port", "sample": sample,
                               "ms_per_sample": ms}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
