"""Multi-GPU correctness on hardware (needs >= 2 CUDA devices; skipped otherwise): the product CLIs launched as 2 NCCL
ranks must write exactly what 1 rank writes, and both must match the reference's golden output.
  - mustache CLI on chromosomes s1 + s2 of BASELINE configs[3] (owners read one chromosome each, 19 blocks are split
    10 / 9, the surplus block's COO travels over all_to_all_single);
  - diff_mustache CLI on BASELINE configs[4] (two 20k-bin maps, 13 block pairs over 2 GPUs)."""
import os
import socket
import sys

import pytest

from mustache_b200 import synth as gen
from tests import synth
from tests.test_gpu_e2e import _read_tsv
from tests.test_gpu_configs import _same_rows

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = synth.GOLDEN


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _rank_main(rank, world, port, which, argv):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    if which == "mustache":
        from mustache_b200 import mustache as m
    else:
        from mustache_b200 import diff_mustache as m
    m.main(argv)
    dist.barrier()
    dist.destroy_process_group()


def _launch(which, argv, world=2):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, which, argv)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=900)
        assert p.exitcode == 0


@pytest.mark.skipif(not _two_gpus(), reason="needs two CUDA devices")
def test_two_rank_cli_equals_one_rank(tmp_path):
    from mustache_b200 import mustache as mm
    names = ["s1", "s2"]
    path = str(tmp_path / "contacts.txt")
    for k, name in enumerate(names):
        spec = {a: b for a, b in gen.CONFIG4[name].items() if a != "res"}
        x, y, c = gen.synthetic_chromosome(**spec)
        gen.write_contact_text(path, name, x, y, c, 5000, mode="w" if k == 0 else "a")
    args = ["-f", path, "-ch"] + names + ["-r", "5kb", "-pt", "0.1", "-st", "0.8"]
    one, two = str(tmp_path / "one.tsv"), str(tmp_path / "two.tsv")
    mm.main(args + ["-o", one])
    _launch("mustache", args + ["-o", two])
    key = lambda r: (r[0], int(r[1]), int(r[4]))
    a, b = sorted(_read_tsv(one), key=key), sorted(_read_tsv(two), key=key)
    assert a == b                                           # identical strings, FDR included
    ref = sorted([r for r in _read_tsv(os.path.join(G, "cfg4_loops.tsv")) if r[0] in names], key=key)
    _same_rows(b, ref)


@pytest.mark.skipif(not _two_gpus(), reason="needs two CUDA devices")
def test_two_rank_differential_cli(tmp_path):
    spec = gen.CONFIG5
    A, B = gen.config5_maps(**spec)
    fa = gen.write_contact_text(str(tmp_path / "mapA.txt"), "chrD", *A, spec["res"])
    fb = gen.write_contact_text(str(tmp_path / "mapB.txt"), "chrD", *B, spec["res"])
    out = str(tmp_path / "diff2")
    _launch("diff", ["-f1", fa, "-f2", fb, "-ch", "chrD", "-r", "5kb", "-pt", "0.05", "-pt2", "0.1", "-st", "0.8", "-o", out])
    for suf in ("loop1", "loop2", "diffloop1", "diffloop2"):
        _same_rows(_read_tsv(out + "." + suf), _read_tsv(os.path.join(G, "cfg5_%s.tsv" % suf)))
