#!/usr/bin/env python3
"""Drop-in mirror of the reference's mustache/mustache.py interface with the scale-space loop on the B200.

Same names, argument meaning and output as the reference (ay-lab/mustache v1.3.3):
  parse_args / main            mustache.py:52-178, 963-1111   flags and TSV columns unchanged
  regulator                    mustache.py:853-942            read -> normalise -> tile; the per-block process fan-out
                                                              (mustache.py:913-937) becomes one batched GPU dispatch
  process_block / mustache     mustache.py:945-960, 697-850   scale-space loop on the engine, post-processing on the host
The hot path has no CPU fallback: without the CUDA library or a GPU these functions raise.

Multi-GPU: when launched under torchrun (WORLD_SIZE > 1) every chromosome is read and normalised by one owner rank, the
blocks of all chromosomes are spread evenly over the ranks (blocks are independent, mustache.py:697-850; sharding.py),
each rank post-processes what it computed and rank 0 gathers the calls and writes the TSV.
"""
import argparse
import os
import sys
import time

import numpy as np

from . import postprocess, readers, sharding, tiler
from .normalize import normalize, normalize_sparse

_ENGINES = {}
_PROGRAM_KEY = {}


def get_engine(device=None):
    """One engine per (process, device).  Device: MUSTACHE_GPU env var, else LOCAL_RANK, else 0."""
    from .engine import ScaleSpaceEngine
    if device is None:
        device = int(os.environ.get("MUSTACHE_GPU", os.environ.get("LOCAL_RANK", "0")))
    if device not in _ENGINES:
        _ENGINES[device] = ScaleSpaceEngine(device)
        if os.environ.get("MUSTACHE_FAST", "0") == "1":          # opt-in FMA arithmetic (include/mustache_b200.h)
            _ENGINES[device].set_arithmetic(True)
    return _ENGINES[device]


def _set_octaves(eng, octave_values):
    key = tuple(float(o) for o in octave_values)
    if _PROGRAM_KEY.get(id(eng)) != key:
        eng.set_octaves(key)
        _PROGRAM_KEY[id(eng)] = key


def parseBP(s):
    """mustache.py:29-49: '5kb' -> 5000, '2mb' -> 2000000, plain digits -> int, anything else -> False."""
    if not s:
        return False
    if s.isnumeric():
        return int(s)
    s = s.lower()
    for suffix, mult in (("kb", 1000), ("mb", 1000000)):
        if suffix in s:
            head = s.split(suffix)[0]
            return int(head) * mult if head.isnumeric() else False
    return False


def parse_args(args):
    """Flag-for-flag the reference's parser (mustache.py:52-178), including its defaults (-pt 0.2, -st 0.88, -oc 2)."""
    p = argparse.ArgumentParser(description="Check the help flag")
    p.add_argument("-f", "--file", dest="f_path", help="REQUIRED: Contact map", required=False)
    p.add_argument("-d", "--distance", dest="distFilter",
                   help="REQUIRED: Maximum distance (in bp) allowed between loop loci", required=False)
    p.add_argument("-o", "--outfile", dest="outdir", help="REQUIRED: Name of the output file.", required=True)
    p.add_argument("-r", "--resolution", dest="resolution", help="REQUIRED: Resolution used for the contact maps",
                   required=True)
    p.add_argument("-bed", "--bed", dest="bed", help="BED file for HiC-Pro type input", default="", required=False)
    p.add_argument("-m", "--matrix", dest="mat", help="MATRIX file for HiC-Pro type input", default="", required=False)
    p.add_argument("-b", "--biases", dest="biasfile",
                   help="RECOMMENDED: biases calculated by ICE or KR norm for each locus", required=False)
    p.add_argument("-cz", "--chromosomeSize", default="", dest="chrSize_file",
                   help="RECOMMENDED: .hic corresponding chromosome size file.", required=False)
    p.add_argument("-norm", "--normalization", default=False, dest="norm_method",
                   help="RECOMMENDED: Hi-C normalization method (KR, VC,...).", required=False)
    p.add_argument("-st", "--sparsityThreshold", dest="st", type=float, default=0.88,
                   help="OPTIONAL: sparsity threshold. Default value is 0.88.", required=False)
    p.add_argument("-pt", "--pThreshold", dest="pt", type=float, default=0.2,
                   help="OPTIONAL: P-value threshold for the results in the final output. Default is 0.2", required=False)
    p.add_argument("-sz", "--sigmaZero", dest="s_z", type=float, default=1.6,
                   help="OPTIONAL: sigma0 value for the method. DEFAULT is 1.6.", required=False)
    p.add_argument("-oc", "--octaves", dest="octaves", default=2, type=int,
                   help="OPTIONAL: Octave count for the method. DEFAULT is 2.", required=False)
    p.add_argument("-i", "--iterations", dest="s", default=10, type=int,
                   help="OPTIONAL: iteration count (ignored by the reference: s = 10 is hard-coded, mustache.py:711)",
                   required=False)
    p.add_argument("-p", "--processes", dest="nprocesses", default=4, type=int,
                   help="OPTIONAL: accepted for compatibility; blocks are batched on the GPU instead", required=False)
    p.add_argument("-ch", "--chromosome", dest="chromosome", nargs="+",
                   help="REQUIRED: Specify which chromosome to run the program for.", default="n", required=False)
    p.add_argument("-ch2", "--chromosome2", dest="chromosome2", nargs="+",
                   help="Optional: second chromosome for interchromosomal analysis (unsupported, as in the reference).",
                   default="n", required=False)
    p.add_argument("-v", "--verbose", dest="verbose", type=bool, default=True, help="OPTIONAL: Verbosity of the program",
                   required=False)
    return p.parse_args(args)


# ------------------------------------------------------------------------------------------------------------------
# block level
# ------------------------------------------------------------------------------------------------------------------
def _loops_from_records(n, dpx, start, mask, rec, st, pt):
    """Everything after mustache.py:772 for one block."""
    mr, mc, mv = mask
    if rec is None or rec["nz_count"] < 50:                  # mustache.py:701-702
        return []
    loops, _ = postprocess.call_loops(n, dpx, start, mr, mc, mv, rec["rows"], rec["cols"], rec["p"], rec["sigma"], st, pt)
    return loops


def mustache(c, chromosome, chromosome2, res, pval_weights, start, end, mask_size, distance_in_px, octave_values, st,
             pt):
    """Same contract as the reference's mustache() (mustache.py:697-850) for one dense tile `c`.

    `c` is modified in place the way the reference modifies it (the 2-fills, mustache.py:703-706).
    """
    if chromosome != chromosome2:
        raise NotImplementedError("inter-chromosomal tiles: the reference path is broken (mustache.py:939-942); unsupported")
    from . import blockrun
    n = c.shape[0]
    d = np.subtract.outer(np.arange(n), np.arange(c.shape[1])) * -1
    nzmask = (c != 0) & (d >= 4)
    if (nzmask & (d > distance_in_px + 1)).any():
        # regulator() never builds such a tile (the readers keep |j - i| <= dpx + 1, mustache.py:264); the engine stores the
        # band only, so refuse loudly instead of scoring a different mask than the reference would
        raise ValueError("tile holds contacts beyond distance_in_px + 1 diagonals; not supported by the banded engine")
    mr, mc = np.nonzero(nzmask)
    mv = c[mr, mc]
    if len(mr) < 50:
        return []
    eng = get_engine()
    _set_octaves(eng, octave_values)
    task = blockrun.BlockTask(0, 0, [(mr, mc, np.ascontiguousarray(mv, dtype=np.float64))])
    (_, cands), = blockrun.run_batches(eng, [task], n, distance_in_px, select=(pt, st))
    assert cands[0]["nz_count"] == len(mr)
    c[d <= 4] = 2
    c[d >= distance_in_px + 1] = 2
    return postprocess.call_loops_from_candidates(n, distance_in_px, start, mr, mc, mv, cands[0])


def process_block(i, start, end, overlap_size, loops, o):
    """Overlap de-duplication of mustache.py:945-960 applied to the loops of block i."""
    ms = tiler.block_mask_size(i, start, end, overlap_size)
    for loop in loops:
        if tiler.keep_after_overlap(loop, start[i], ms):
            o.append([loop[0], loop[1], loop[2], loop[3]])


def _dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend)
    return rank, world


def _block_calls(dpx, st, pt):
    """Host half of the post-processing of one block, on the rank that computed it: the device delivered the candidates
    that passed BH, `o < pt` and the sparsity filter (mustache.py:774-811); enrichment filter and clustering here
    (mustache.py:816-848)."""
    def fn(task, cands, chunk, start):
        mr, mc, mv = task.maps[0]
        if cands[0]["nz_count"] < 50:                            # mustache.py:701-702
            return []
        return postprocess.call_loops_from_candidates(chunk, dpx, start, mr, mc, mv, cands[0])
    return fn


def call_chromosomes(preps, n_chrom, distance_in_px, octave_values, st, pt, verbose=True, rank=0, world=1, owners=None,
                     timings=None):
    """The block pool of `n_chrom` chromosomes: preps = {chromosome index: dict(maps=[(x, y, v)], n=bins)} for the
    chromosomes this rank read and normalised.  Every block runs on exactly one rank (sharding.balanced_assignment),
    is post-processed there, and rank 0 receives {chromosome index: de-duplicated loop list}."""
    from . import blockrun
    eng = get_engine()
    _set_octaves(eng, octave_values)
    return blockrun.shard_and_call(preps, n_chrom, distance_in_px, 1, eng, _block_calls(distance_in_px, st, pt), rank=rank,
                                   world=world, verbose=verbose, owners=owners, timings=timings, width=4, select=(pt, st))


def call_blocks(x, y, v, n, distance_in_px, octave_values, st, pt, verbose=True, rank=0, world=1, device=None,
                timings=None):
    """One normalised chromosome (held by rank 0; what other ranks pass is ignored): tile it, run every block on some
    rank, post-process, return the de-duplicated loop list (rank 0; other ranks get [])."""
    preps = {0: dict(maps=[(np.asarray(x), np.asarray(y), np.asarray(v))], n=int(n))} if rank == 0 else {}
    out = call_chromosomes(preps, 1, distance_in_px, octave_values, st, pt, verbose=verbose, rank=rank, world=world,
                           owners=[0], timings=timings)
    return out.get(0, [])


def prepare_chromosome(f, norm_method, CHRM_SIZE, res, distance_filter, bias, chromosome, chromosome2, verbose=True):
    """Read and normalise one chromosome (mustache.py:877-895): returns dict(maps=[(x, y, v)], n=bins) or None."""
    if verbose:
        print("Reading contact map...")
    if f.endswith(".hic"):
        got = readers.read_hic(f, norm_method, CHRM_SIZE, distance_filter, chromosome, chromosome2, res)
    elif f.endswith(".cool") or f.endswith(".mcool"):
        got = readers.read_cool(f, distance_filter, chromosome, chromosome2, norm_method, res)
    else:
        got = readers.read_text(f, distance_filter, bias, chromosome, res)
    if got is None:
        return None
    x, y, v = got
    if len(v) == 0:
        return None
    if verbose:
        print("Normalizing contact map...")
    dpx = tiler.distance_in_px(distance_filter, res)
    n = int(max(np.max(x), np.max(y)) + 1)
    normalize(x, y, v, res, dpx, eng=get_engine(), biased=bool(bias))      # in place
    return dict(maps=[(np.asarray(x), np.asarray(y), np.asarray(v))], n=n)


def regulator(f, norm_method, CHRM_SIZE, outdir, bed="", res=5000, sigma0=1.6, s=10, pt=0.1, st=0.88, octaves=2,
              verbose=True, nprocesses=4, distance_filter=2000000, bias=False, chromosome="n", chromosome2=None):
    """mustache.py:853-942 with the block fan-out on the GPU(s).  Returns [[x, y, fdr, scale], ...] in bin units."""
    if not chromosome2 or chromosome2 == "n":
        chromosome2 = chromosome
    if chromosome != chromosome2:
        print("Interchromosomal analysis is only supported for .hic and .cool input formats.")
        raise FileNotFoundError
    octave_values = [sigma0 * (2 ** i) for i in range(octaves)]
    rank, world = _dist_env()
    preps = {}
    if rank == 0:                      # one reader per chromosome; the blocks are spread over all ranks afterwards
        prep = prepare_chromosome(f, norm_method, CHRM_SIZE, res, distance_filter, bias, chromosome, chromosome2, verbose)
        if prep is not None:
            preps[0] = prep
    if verbose:
        print("Loop calling...")
    dpx = tiler.distance_in_px(distance_filter, res)
    out = call_chromosomes(preps, 1, dpx, octave_values, st, pt, verbose=verbose, rank=rank, world=world, owners=[0])
    return out.get(0, [])


def resolve_distance(dist_arg, res, cap=10000):
    """Distance-limit policy of mustache.py:996-1015 (`cap` is 2000 in diff_mustache.py:773-778)."""
    dist = parseBP(dist_arg)
    if not dist:
        if 200 * res >= 2000000:
            dist = 200 * res
            print("The distance limit is set to {}bp".format(200 * res))
        elif 2000 * res <= 2000000:
            dist = 2000 * res
            print("The distance limit is set to {}bp".format(2000 * res))
        else:
            dist = 2000000
            print("The distance limit is set to 2Mbp")
    elif dist < 200 * res:
        print("The distance limit is set to {}bp".format(200 * res))
        dist = 200 * res
    elif dist > cap * res:
        print("The distance limit is set to {}bp".format(cap * res))
        dist = cap * res
    elif dist > 10000000:
        dist = 10000000
        print("The distance limit is set to 10Mbp")
    return dist


HEADER = "BIN1_CHR\tBIN1_START\tBIN1_END\tBIN2_CHROMOSOME\tBIN2_START\tBIN2_END\tFDR\tDETECTION_SCALE\n"


def format_row(chromosome, chromosome2, loop, res):
    """One TSV row exactly as mustache.py:1098-1103 writes it (str() of numpy scalars)."""
    x, y = np.int64(loop[0]), np.int64(loop[1])
    return (str(chromosome) + "\t" + str(x * res) + "\t" + str((x + 1) * res) + "\t" + str(chromosome2) + "\t"
            + str(y * res) + "\t" + str((y + 1) * res) + "\t" + str(np.float64(loop[2])) + "\t" + str(np.float64(loop[3]))
            + "\n")


def main(argv=None):
    start_time = time.time()
    args = parse_args(sys.argv[1:] if argv is None else argv)
    rank, world = _dist_env()
    quiet = rank != 0
    if not quiet:
        print("\n")
    f = args.f_path
    if args.bed and args.mat:
        f = args.mat
    if not f or not os.path.exists(f):
        print("Error: Couldn't find the specified contact files")
        return
    res = parseBP(args.resolution)
    if not res:
        print("Error: Invalid resolution")
        return
    if not args.chromosome or args.chromosome == "n":
        if f.endswith(".cool") or f.endswith(".mcool") or f.endswith(".hic"):
            print("Error: chromosome enumeration from .hic/.cool needs hicstraw/cooler (absent); pass -ch explicitly")
            return
        print("Error: Please enter the chromosome name.")
        return
    distFilter = resolve_distance(args.distFilter, res)
    chr_list = list(args.chromosome)
    if (args.chromosome2 and args.chromosome2 != "n") and len(chr_list) != len(args.chromosome2):
        print("Error: the same number of chromosome1 and chromosome2 should be provided.")
        return
    chr_list2 = list(args.chromosome2) if isinstance(args.chromosome2, list) else list(chr_list)
    chr_sizes = False
    if args.chrSize_file:
        import pandas as pd
        csz = pd.read_csv(args.chrSize_file, header=None, sep="\t")
        chr_sizes = {"chr" + str(csz.iloc[i, 0]).replace("chr", ""): csz.iloc[i, 1] for i in range(csz.shape[0])}
    biasf = False
    if args.biasfile:
        if os.path.exists(args.biasfile):
            biasf = args.biasfile
        else:
            print("Error: Couldn't find specified bias file")
            return
    for c1, c2 in zip(chr_list, chr_list2):
        if c1 != c2:
            print("Interchromosomal analysis is only supported for .hic and .cool input formats.")
            raise FileNotFoundError
    octave_values = [args.s_z * (2 ** k) for k in range(args.octaves)]
    dpx = tiler.distance_in_px(distFilter, res)
    verbose = args.verbose and not quiet
    # The reference walks the chromosomes one after the other (mustache.py:1057).  Here a round of chromosomes is read
    # and normalised by their owner ranks, and the blocks of the whole round form one pool spread over all GPUs.
    per_round = max(8, 2 * world)
    wrote_header = False
    for r0 in range(0, len(chr_list), per_round):
        names = chr_list[r0:r0 + per_round]
        sizes = [chr_sizes["chr" + str(c).replace("chr", "")] for c in names] if chr_sizes else None
        owners = sharding.chromosome_owners(len(names), world, sizes)
        preps = {}
        for k, chromosome in enumerate(names):
            if owners[k] != rank:
                continue
            CHRM_SIZE = sizes[k] if sizes else False
            prep = prepare_chromosome(f, args.norm_method, CHRM_SIZE, res, distFilter, biasf, chromosome, chromosome, verbose)
            if prep is not None:
                preps[k] = prep
        if verbose:
            print("Loop calling...")
        found = call_chromosomes(preps, len(names), dpx, octave_values, args.st, args.pt, verbose=verbose, rank=rank,
                                 world=world, owners=owners)
        if quiet:
            continue
        for k, chromosome in enumerate(names):
            o = found.get(k, [])
            if not wrote_header:
                with open(args.outdir, "w") as out_file:
                    out_file.write(HEADER)
                wrote_header = True
            print("{0} loops found for chrmosome={1}, fdr<{2} in {3}sec".format(len(o), chromosome, args.pt,
                                                                                "%.2f" % (time.time() - start_time)))
            if o:
                with open(args.outdir, "a") as out_file:
                    for loop in o:
                        out_file.write(format_row(chromosome, chromosome, loop, res))
            start_time = time.time()


if __name__ == "__main__":
    main()
