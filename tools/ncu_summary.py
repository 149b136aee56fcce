#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`): headline metrics per kernel and, for one kernel, the
instruction / stall-sample split between barrier-delimited phases of the SASS.  Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --traffic CONFIG profiles/traffic_r01.json
(the second form records dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel under CONFIG; bench.py
reports it as roofline.traffic)"""
import json
import os
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_uniform.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    if kre == "--traffic":
        cfg, out = sys.argv[3], sys.argv[4]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        acc = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
            tot = sum(float(d[m].replace(",", "")) * scale[units[hdr.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            acc.setdefault(name, []).append(tot)
        data = json.load(open(out)) if os.path.exists(out) else {}
        data[cfg] = {k: sum(v) / len(v) for k, v in acc.items()}
        json.dump(data, open(out, "w"), indent=1, sort_keys=True)
        print(json.dumps(data[cfg]))
        return
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  grid %s block %s" % (d["Kernel Name"].split("(")[0], d.get("Grid Size"), d.get("Block Size")))
        for w in WANT:
            if w in d:
                print("   %-66s %s %s" % (w, d[w], units[hdr.index(w)]))
        stalls = {k: float(v.replace(",", "")) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")}
        top = sorted(stalls.items(), key=lambda x: -x[1])[:6]
        print("   top stalls (warps per issue):", [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(v, 2)) for k, v in top])
    if not kre:
        return
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre]))))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    h = rows[hi]
    isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = [(r[isrc].strip(), int(r[iex]), int(r[ismp])) for r in rows[hi + 1:] if len(r) > iex and r[iex].isdigit()]
    tot, tots = sum(d[1] for d in data), sum(d[2] for d in data)
    seg, segs = 0, collections.defaultdict(lambda: [0, 0, collections.Counter(), 0])
    for s, ex, sm in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
        segs[seg][0] += ex; segs[seg][1] += sm; segs[seg][2][op] += ex; segs[seg][3] += 1
        if "BAR.SYNC" in s or "SYNCS.PHASECHK" in s:
            seg += 1
    print("-- SASS phases of %s (split at BAR.SYNC / mbarrier waits): static instrs, %% of executed instrs, %% of stall samples, top opcodes" % kre)
    for k, v in segs.items():
        if v[0] * 200 < tot and v[1] * 200 < tots:
            continue
        print("   phase %d: n=%d inst %.1f%% samples %.1f%% %s" % (k, v[3], 100 * v[0] / tot, 100 * v[1] / max(tots, 1),
                                                                 [(o, round(100 * c / tot, 1)) for o, c in v[2].most_common(8)]))
    ops = collections.Counter()
    for s, ex, sm in data:
        ops[re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]] += ex
    print("   executed by opcode:", [(o, round(100 * c / tot, 1)) for o, c in ops.most_common(14)])


if __name__ == "__main__":
    main()
