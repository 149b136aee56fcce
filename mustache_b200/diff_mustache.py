#!/usr/bin/env python3
"""Drop-in mirror of the reference's mustache/diff_mustache.py interface (two-map differential loop calling) with the
three scale-space stacks on the B200.

  diff_mustache   diff_mustache.py:260-569   per block: both maps scored, difference stack, BH per map, filters,
                                             clustering, differential selection (pair < pt2 and v_self > v_other)
  regulator       diff_mustache.py:572-690   reads both maps, normalises each, tiles, batches block pairs on the GPU
  main            diff_mustache.py:720-906   -f1 -f2 -b1 -b2 -pt -pt2 ...; writes .loop1 .loop2 .diffloop1 .diffloop2
Reference quirks kept for parity: #13 `-b1` is never applied to map 1 for text input (the CLI
reads `args.biasfile1` into `biasf` instead of `biasf1`), #15 pt2 thresholds the raw two-sided normal p, #16 explicit
-d is capped at 2000*res.
"""
import argparse
import os
import sys
import time

import numpy as np

from . import postprocess, readers, tiler
from .mustache import HEADER, _dist_env, _set_octaves, format_row, get_engine, parseBP, resolve_distance
from .normalize import normalize

_DIFF_KEY = {}


def _set_octaves_diff(eng, octave_values):
    from . import mustache as _m
    key = tuple(float(o) for o in octave_values)
    # the difference chain must belong to the main chain the engine holds NOW (mustache() may have re-programmed it)
    if _DIFF_KEY.get(id(eng)) != key or _m._PROGRAM_KEY.get(id(eng)) != key:
        eng.set_octaves(key, differential=True)
        _DIFF_KEY[id(eng)] = key
        _m._PROGRAM_KEY[id(eng)] = key


def _dense_lookup(index, values, rows, cols, on_mask_default, off_mask):
    pos = index.lookup(rows, cols)
    return np.where(pos >= 0, values[np.maximum(pos, 0)], off_mask)


def select_differential(n, dpx, start, masks, recs, st, pt, pt2):
    """diff_mustache.py:428-569 for one block pair.  masks/recs: [(rows, cols, vals)] and engine records per map."""
    if any(len(m[0]) < 50 for m in masks):                              # diff_mustache.py:266-267
        return [], [], [], []
    if any(len(m[0]) < postprocess.MIN_MASK_FOR_BH for m in masks):      # diff_mustache.py:430-431
        return [], [], [], []
    outs, auxs = [], []
    for m, r in zip(masks, recs):
        loops, aux = postprocess.call_loops(n, dpx, start, m[0], m[1], m[2], r["rows"], r["cols"], r["p"], r["sigma"],
                                            st, pt, candidate_order="rowmajor", partial=True)
        outs.append(loops)
        auxs.append(aux)
    # the reference bails out of the whole block when either map has no candidate left after the filters
    if any(a.get("empty_after_filters") for a in auxs):                  # diff_mustache.py:507-508, 519-520, 526-527
        return [], [], [], []
    dense = []
    for m, r, a in zip(masks, recs, auxs):
        idx = a["index"]
        pos = idx.lookup(r["rows"], r["cols"])
        pair = np.full(idx.keys.size, 2.0)                               # pPair initialised to 2 (diff_mustache.py:290)
        vall = np.zeros(idx.keys.size)                                   # vAll initialised to 0 (diff_mustache.py:292)
        pair[pos] = r["pair"]
        vall[pos] = r["v"]
        dense.append((idx, pair, vall))

    def lookup(which, vals, rows, cols):
        idx = dense[which][0]
        pos = idx.lookup(rows, cols)
        return np.where(pos >= 0, vals[np.maximum(pos, 0)], 1.0)        # np.ones_like off the mask (:447-453)

    diffs = []
    for me, other in ((0, 1), (1, 0)):
        keep = []
        for loop in outs[me]:
            rr, cc = [loop[0] - start], [loop[1] - start]
            pr = lookup(me, dense[me][1], rr, cc)[0]
            v_self = lookup(me, dense[me][2], rr, cc)[0]
            v_other = lookup(other, dense[other][2], rr, cc)[0]
            if pr < pt2 and v_self > v_other:                            # diff_mustache.py:567-568
                keep.append(loop)
        diffs.append(keep)
    return outs[0], diffs[0], outs[1], diffs[1]


def select_differential_from_candidates(n, dpx, start, masks, cands, pt2):
    """diff_mustache.py:428-569 for one block pair when the device already did BH, `o < pt` and the sparsity filter per map
    (mb200_select_candidates after mb200_run_differential): cands = the two entries of candidates_batch(pair=True)."""
    if any(len(m[0]) < 50 for m in masks):                              # diff_mustache.py:266-267
        return [], [], [], []
    if any(len(m[0]) < postprocess.MIN_MASK_FOR_BH for m in masks):      # diff_mustache.py:430-431
        return [], [], [], []
    outs = []
    for m, c in zip(masks, cands):
        loops, vals, emptied = postprocess.call_loops_from_candidates(n, dpx, start, m[0], m[1], m[2], c,
                                                                      extra=("pair9", "vself9", "vother9"))
        if emptied:                                                      # diff_mustache.py:507-508, 519-520, 526-527
            return [], [], [], []
        outs.append((loops, vals))
    diffs = [[l for l, v in zip(loops, vals) if v["pair9"] < pt2 and v["vself9"] > v["vother9"]]     # diff_mustache.py:567-568
             for loops, vals in outs]
    return outs[0][0], diffs[0], outs[1][0], diffs[1]


def diff_mustache(c1, c2, chromosome, chromosome2, res, start, end, mask_size, distance_in_px, octave_values, st, pt, pt2):
    """Same contract as the reference's diff_mustache() (diff_mustache.py:260-569) for one pair of dense tiles."""
    if chromosome != chromosome2:
        raise NotImplementedError("inter-chromosomal tiles are not supported (broken in the reference)")
    n = c1.shape[0]
    d = np.subtract.outer(np.arange(n), np.arange(n)) * -1
    masks = []
    for c in (c1, c2):
        nzmask = (c != 0) & (d >= 4)
        if (nzmask & (d > distance_in_px + 1)).any():
            raise ValueError("tile holds contacts beyond distance_in_px + 1 diagonals; not supported by the banded engine")
        r, cc = np.nonzero(nzmask)
        masks.append((r, cc, c[r, cc]))
    if any(len(m[0]) < 50 for m in masks):
        return [], [], [], []
    eng = get_engine()
    _set_octaves_diff(eng, octave_values)
    from . import blockrun
    task = blockrun.BlockTask(0, 0, [(m[0], m[1], np.ascontiguousarray(m[2], dtype=np.float64)) for m in masks])
    (_, cands), = blockrun.run_batches(eng, [task], n, distance_in_px, differential=True, select=(pt, st))
    assert [r["nz_count"] for r in cands] == [len(m[0]) for m in masks]
    for c in (c1, c2):
        c[d <= 4] = 2
        c[d >= distance_in_px + 1] = 2
    return select_differential_from_candidates(n, distance_in_px, start, masks, cands, pt2)


def _pair_calls(dpx, st, pt, pt2):
    """Selection of one block pair where it was computed (diff_mustache.py:428-569): calls carry the tag of the output
    they belong to, 1/2/3/4 = loops1 / diffloops1 / loops2 / diffloops2 (diff_mustache.py:704-715)."""
    def fn(task, cands, chunk, start):
        out = []
        for tag, loops in zip((1, 2, 3, 4), select_differential_from_candidates(chunk, dpx, start, task.maps, cands, pt2)):
            out += [[l[0], l[1], l[2], l[3], tag] for l in loops]
        return out
    return fn


def call_chromosome_pairs(preps, n_chrom, dpx, octave_values, st, pt, pt2, verbose=True, rank=0, world=1, owners=None):
    """Block pool of the differential run: preps = {chromosome index: dict(maps=[(x, y, v) map 1, (x, y, v) map 2], n)}
    for the chromosomes this rank read.  Both tiles of a block pair stay on one GPU (the difference stack needs them and
    their common mask).  Rank 0 receives {chromosome index: [[x, y, fdr, scale, tag]]}."""
    from . import blockrun
    eng = get_engine()
    _set_octaves_diff(eng, octave_values)
    return blockrun.shard_and_call(preps, n_chrom, dpx, 2, eng, _pair_calls(dpx, st, pt, pt2), rank=rank, world=world,
                                   verbose=verbose, owners=owners, width=5, select=(pt, st))


def call_block_pairs(xyv1, xyv2, n, dpx, octave_values, st, pt, pt2, verbose=True, rank=0, world=1):
    """Two normalised maps of one chromosome (held by rank 0).  Returns tagged loops [[x, y, fdr, scale, tag]] on rank 0."""
    preps = {0: dict(maps=[tuple(np.asarray(a) for a in xyv1), tuple(np.asarray(a) for a in xyv2)], n=int(n))} if rank == 0 else {}
    out = call_chromosome_pairs(preps, 1, dpx, octave_values, st, pt, pt2, verbose=verbose, rank=rank, world=world, owners=[0])
    return [[l[0], l[1], l[2], l[3], int(l[4])] for l in out.get(0, [])]


def prepare_pair(f1, f2, norm_method, CHRM_SIZE, res, distance_filter, bias1, bias2, chromosome, chromosome2, verbose=True):
    """Read and normalise both maps of one chromosome (diff_mustache.py:602-635)."""
    if verbose:
        print("Reading contact map...")
    maps = []
    for f, b in ((f1, bias1), (f2, bias2)):
        if f.endswith(".hic"):
            got = readers.read_hic(f, norm_method, CHRM_SIZE, distance_filter, chromosome, chromosome2, res)
        elif f.endswith(".cool") or f.endswith(".mcool"):
            got = readers.read_cool(f, distance_filter, chromosome, chromosome2, norm_method, res)
        else:
            got = readers.read_text(f, distance_filter, b, chromosome, res)
        if got is None or len(got[2]) == 0:
            return None
        maps.append([np.asarray(a) for a in got])
    if verbose:
        print("Normalizing contact map...")
    dpx = tiler.distance_in_px(distance_filter, res)
    n = int(max(max(m[0].max(), m[1].max()) + 1 for m in maps))
    for m, b in zip(maps, (bias1, bias2)):
        normalize(m[0], m[1], m[2], res, dpx, eng=get_engine(), biased=bool(b))
    return dict(maps=[tuple(m) for m in maps], n=n)


def regulator(f1, f2, norm_method, CHRM_SIZE, outdir, bed1="", bed2="", res=5000, sigma0=1.6, s=10, pt=0.1, pt2=0.1,
              st=0.88, octaves=2, verbose=True, nprocesses=4, distance_filter=2000000, bias1=False, bias2=False,
              chromosome="n", chromosome2=None):
    if not chromosome2 or chromosome2 == "n":
        chromosome2 = chromosome
    if chromosome != chromosome2:
        print("Interchromosomal analysis is only supported for .hic and .cool input formats.")
        raise FileNotFoundError
    octave_values = [sigma0 * (2 ** i) for i in range(octaves)]
    rank, world = _dist_env()
    preps = {}
    if rank == 0:
        prep = prepare_pair(f1, f2, norm_method, CHRM_SIZE, res, distance_filter, bias1, bias2, chromosome, chromosome2, verbose)
        if prep is not None:
            preps[0] = prep
    if verbose:
        print("Loop calling...")
    dpx = tiler.distance_in_px(distance_filter, res)
    out = call_chromosome_pairs(preps, 1, dpx, octave_values, st, pt, pt2, verbose=verbose, rank=rank, world=world, owners=[0])
    return [[l[0], l[1], l[2], l[3], int(l[4])] for l in out.get(0, [])]


def parse_args(args):
    """diff_mustache.py:29-180."""
    p = argparse.ArgumentParser(description="Check the help flag")
    p.add_argument("-f1", "--file1", dest="f_path1", required=False)
    p.add_argument("-f2", "--file2", dest="f_path2", required=False)
    p.add_argument("-d", "--distance", dest="distFilter", required=False)
    p.add_argument("-o", "--outfile", dest="outdir", required=True)
    p.add_argument("-r", "--resolution", dest="resolution", required=True)
    p.add_argument("-bed1", "--bed1", dest="bed1", default="", required=False)
    p.add_argument("-bed2", "--bed2", dest="bed2", default="", required=False)
    p.add_argument("-m1", "--matrix1", dest="mat1", default="", required=False)
    p.add_argument("-m2", "--matrix2", dest="mat2", default="", required=False)
    p.add_argument("-b1", "--biases1", dest="biasfile1", required=False)
    p.add_argument("-b2", "--biases2", dest="biasfile2", required=False)
    p.add_argument("-cz", "--chromosomeSize", default="", dest="chrSize_file", required=False)
    p.add_argument("-norm", "--normalization", default=False, dest="norm_method", required=False)
    p.add_argument("-st", "--sparsityThreshold", dest="st", type=float, default=0.88, required=False)
    p.add_argument("-pt", "--pThreshold", dest="pt", type=float, default=0.2, required=False)
    p.add_argument("-pt2", "--pThreshold2", dest="pt2", type=float, default=0.1, required=False)
    p.add_argument("-sz", "--sigmaZero", dest="s_z", type=float, default=1.6, required=False)
    p.add_argument("-oc", "--octaves", dest="octaves", default=2, type=int, required=False)
    p.add_argument("-i", "--iterations", dest="s", default=10, type=int, required=False)
    p.add_argument("-p", "--processes", dest="nprocesses", default=4, type=int, required=False)
    p.add_argument("-ch", "--chromosome", dest="chromosome", nargs="+", default="n", required=False)
    p.add_argument("-ch2", "--chromosome2", dest="chromosome2", nargs="+", default="n", required=False)
    p.add_argument("-v", "--verbose", dest="verbose", type=bool, default=True, required=False)
    return p.parse_args(args)


def main(argv=None):
    start_time = time.time()
    args = parse_args(sys.argv[1:] if argv is None else argv)
    rank, world = _dist_env()
    quiet = rank != 0
    f1, f2 = args.f_path1, args.f_path2
    if args.bed1 and args.mat1:
        f1 = args.mat1
    if args.bed2 and args.mat2:
        f2 = args.mat2
    if not f1 or not f2 or not os.path.exists(f1) or not os.path.exists(f2):
        print("Error: Couldn't find the specified contact files")
        return
    res = parseBP(args.resolution)
    if not res:
        print("Error: Invalid resolution")
        return
    if not args.chromosome or args.chromosome == "n":
        print("Error: Please enter the chromosome name.")
        return
    distFilter = resolve_distance(args.distFilter, res, cap=2000)       # diff_mustache.py:770-778 (quirk #16)
    chr_list = list(args.chromosome)
    chr_list2 = list(args.chromosome2) if isinstance(args.chromosome2, list) else list(chr_list)
    first = True
    for chromosome, chromosome2 in zip(chr_list, chr_list2):
        # quirk #13 (diff_mustache.py:824-827, 850): `biasf = args.biasfile1` -- biasf1 stays False, so map 1 is never
        # bias-corrected for text input; only -b2 takes effect.
        biasf1, biasf2 = False, False
        if args.biasfile1 and not os.path.exists(args.biasfile1):
            print("Error: Couldn't find specified bias file1")
            return
        if args.biasfile2:
            if os.path.exists(args.biasfile2):
                biasf2 = args.biasfile2
            else:
                print("Error: Couldn't find specified bias file2")
                return
        o = regulator(f1, f2, args.norm_method, False, args.outdir, bed1=args.bed1, bed2=args.bed2, res=res, sigma0=args.s_z,
                      s=args.s, verbose=args.verbose and not quiet, pt=args.pt, pt2=args.pt2, st=args.st,
                      distance_filter=distFilter, nprocesses=args.nprocesses, bias1=biasf1, bias2=biasf2,
                      chromosome=chromosome, chromosome2=chromosome2, octaves=args.octaves)
        if quiet:
            continue
        names = {1: ".loop1", 2: ".diffloop1", 3: ".loop2", 4: ".diffloop2"}
        if first:
            for suf in names.values():
                with open(args.outdir + suf, "w") as fh:
                    fh.write(HEADER)
            first = False
        counts = {t: 0 for t in names}
        for loop in o:
            counts[loop[4]] += 1
            with open(args.outdir + names[loop[4]], "a") as fh:
                fh.write(format_row(chromosome, chromosome2, loop, res))
        print(f"({counts[1]},{counts[3]}) loops and ({counts[2]},{counts[4]}) differential-loops found in "
              f"chrmosome={chromosome} for detection-fdr<{args.pt} and difference-fdr<{args.pt2} in "
              f"{time.time() - start_time:.2f}sec")
        start_time = time.time()


if __name__ == "__main__":
    main()
