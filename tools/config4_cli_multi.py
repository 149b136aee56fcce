#!/usr/bin/env python3
"""BASELINE configs[3] end to end through the product CLI on N GPUs: the 8 synthetic chromosomes (10k..80k bins at 5 kb) in
ONE contact file, `python -m mustache_b200 -f all.txt -ch s1 .. s8 ...` launched as N NCCL ranks, output compared with
the reference's golden TSV (tests/golden/cfg4_loops.tsv: 368 loops).  Usage: python tools/config4_cli_multi.py [N]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from mustache_b200 import synth as gen
    from tests.test_gpu_configs import _same_rows
    from tests.test_gpu_e2e import _read_tsv
    from tests.test_gpu_multi import _launch
    world = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    tmp = "/tmp/cfg4_cli"
    os.makedirs(tmp, exist_ok=True)
    path = os.path.join(tmp, "all.txt")
    t0 = time.time()
    names = list(gen.CONFIG4)
    for k, name in enumerate(names):
        spec = {a: b for a, b in gen.CONFIG4[name].items() if a != "res"}
        x, y, c = gen.synthetic_chromosome(**spec)
        gen.write_contact_text(path, name, x, y, c, 5000, mode="w" if k == 0 else "a")
    t_gen = time.time() - t0
    out = os.path.join(tmp, "out_%d.tsv" % world)
    argv = ["-f", path, "-ch"] + names + ["-r", "5kb", "-pt", "0.1", "-st", "0.8", "-o", out]
    t0 = time.time()
    if world == 1:
        from mustache_b200 import mustache as mm
        mm.main(argv)
    else:
        _launch("mustache", argv, world=world)
    t_run = time.time() - t0
    key = lambda r: (r[0], int(r[1]), int(r[4]))
    got = sorted(_read_tsv(out), key=key)
    ref = sorted(_read_tsv(os.path.join(ROOT, "tests", "golden", "cfg4_loops.tsv")), key=key)
    _same_rows(got, ref)
    print(json.dumps({"config": "BASELINE configs[3] through the CLI", "n_gpus": world, "chromosomes": len(names), "loops": len(got),
                      "matches_reference_golden": True, "file_mb": os.path.getsize(path) / 1e6, "generate_text_s": t_gen,
                      "cli_wall_s": t_run}))


if __name__ == "__main__":
    main()
