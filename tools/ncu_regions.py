#!/usr/bin/env python3
"""Split the SASS of one kernel of an .ncu-rep into regions (loops found from backward branches) and report, per region,
executed warp instructions, stall samples, FP64 share and the dominant stall reasons.  Usage:
    python tools/ncu_regions.py gpurun_out/prof.ncu-rep kh_kernel [min_len]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    h = rows[hi]
    isrc, iex, ismp, iaddr = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Address")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= iex or not r[iex].isdigit():
            if len(r) > 1 and r[0] == "Kernel Name" and data:
                break                                   # only the first captured launch of the kernel
            continue
        data.append((int(r[iaddr], 16), r[isrc].strip(), int(r[iex]), int(r[ismp]), {c: int(r[i] or 0) for i, c in stall_cols}))
    base = data[0][0]
    data = [(a - base, s, e, m, st) for a, s, e, m, st in data]
    tot_e, tot_m = sum(d[2] for d in data), sum(d[3] for d in data)
    # region boundaries: targets and sources of backward branches (innermost loops win)
    loops = []
    for a, s, e, m, st in data:
        mm = re.search(r"BRA\S*\s+(0x[0-9a-f]+)", s)
        if mm:
            t = int(mm.group(1), 16)
            t = t - base if t >= base else t
            if t < a:
                loops.append((t, a))
    loops.sort(key=lambda x: x[1] - x[0])
    owner = {}
    for li, (lo, hi_) in enumerate(loops):
        for a, *_ in data:
            if lo <= a <= hi_ and a not in owner:
                owner[a] = li
    print("kernel %s: %d SASS instrs, %.3g executed warp-instrs, %d samples" % (kre, len(data), tot_e, tot_m))
    agg = collections.OrderedDict()
    prev = None
    seg = 0
    for a, s, e, m, st in data:
        key = owner.get(a, -1)
        if key != prev:
            seg += 1
            prev = key
        k2 = (seg, key)
        g = agg.setdefault(k2, dict(n=0, e=0, m=0, fp=0, st=collections.Counter(), a0=a, a1=a, ops=collections.Counter()))
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
        g["n"] += 1; g["e"] += e; g["m"] += m; g["a1"] = a
        g["ops"][op] += e
        if op in ("DADD", "DMUL", "DFMA", "DSETP"):
            g["fp"] += e
        for c, v in st.items():
            g["st"][c] += v
    for (seg, key), g in agg.items():
        if g["e"] * 100 < tot_e and g["m"] * 100 < tot_m:
            continue
        top = [(c.replace("stall_", ""), round(100 * v / max(g["m"], 1))) for c, v in g["st"].most_common(4)]
        loop = ("loop %#x-%#x" % loops[key]) if key >= 0 else "straight"
        print("  %#7x-%#7x %-22s n=%4d exec %5.1f%% samples %5.1f%% fp64 %4.0f%% stalls %s ops %s" % (
            g["a0"], g["a1"], loop, g["n"], 100 * g["e"] / tot_e, 100 * g["m"] / tot_m, 100 * g["fp"] / max(g["e"], 1), top,
            [(o, round(100 * c / max(g["e"], 1))) for o, c in g["ops"].most_common(5)]))


if __name__ == "__main__":
    main()
