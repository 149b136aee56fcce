"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/mustache_b200.h declares
(no compute calls without a GPU) and error behaviour without a device.  The world_size-2 gloo tests of the multi-GPU host
path live in tests/test_sharded_blocks.py."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mustache_b200.h")).read()
    return sorted(set(re.findall(r"MB200_API\s+[\w\s\*]+?\b(mb200_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from mustache_b200 import build, engine
    build.build()
    lib = engine.load_library()
    names = _declared_symbols()
    assert len(names) >= 24
    for name in names:
        assert hasattr(lib, name), name
        assert name in engine.SIGNATURES, "ctypes binding missing for " + name
    assert set(engine.SIGNATURES) == set(names)
    assert lib.mb200_abi_version() == 1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mustache_b200 import engine
    with pytest.raises(RuntimeError):
        engine.ScaleSpaceEngine(0)
    lib = engine.load_library()
    h = C.c_void_p()
    assert lib.mb200_create(0, C.byref(h)) < 0 and not h.value
    assert lib.mb200_last_error(None) == b"null engine"
    assert lib.mb200_run(None) < 0


def test_cli_parsers_match_reference_defaults():
    from mustache_b200 import diff_mustache, mustache
    a = mustache.parse_args(["-f", "x", "-r", "5kb", "-o", "out", "-ch", "21"])
    assert (a.pt, a.st, a.s_z, a.octaves, a.s, a.nprocesses) == (0.2, 0.88, 1.6, 2, 10, 4)
    assert a.chromosome == ["21"] and a.chromosome2 == "n"
    assert mustache.parseBP("5kb") == 5000 and mustache.parseBP("2mb") == 2000000 and mustache.parseBP("12") == 12
    assert mustache.parseBP("kb") is False and mustache.parseBP("") is False
    assert mustache.resolve_distance(None, 5000) == 2000000 and mustache.resolve_distance(None, 100000) == 20000000
    assert mustache.resolve_distance(None, 500) == 1000000 and mustache.resolve_distance("100mb", 5000) == 50000000
    assert mustache.resolve_distance("100mb", 5000, cap=2000) == 10000000
    d = diff_mustache.parse_args(["-f1", "a", "-f2", "b", "-r", "5kb", "-o", "o", "-ch", "1", "-pt2", "0.05"])
    assert d.pt2 == 0.05 and d.pt == 0.2
    row = mustache.format_row("21", "21", [3161, 3224, 0.0919438088559645, 2.111212657236631], 5000)
    assert row == "21\t15805000\t15810000\t21\t16120000\t16125000\t0.0919438088559645\t2.111212657236631\n"


def test_kv_plan_is_a_partition_with_the_optimal_cost():
    """The axis-0 kernel's grouping (host-only entry point, runs without a GPU): every step in exactly one group of at
    most 5 (3 for chains that reach radius 24), groups ordered by radius, and the cost equals the dynamic programme bench.py restates for the roofline."""
    import bench
    from mustache_b200 import engine, ladder
    lib = engine.load_library()
    for octs in ([1.6, 3.2], [1.6, 3.2, 6.4, 12.8], [0.7, 1.4, 2.8], [2.3]):
        prog = ladder.build_program(octs)
        radius = np.array([s.radius for s in prog.steps], np.int32)
        grp = np.full(len(radius), -1, np.int32)
        cost = C.c_int64(0)
        st = lib.mb200_kv_plan(len(radius), radius.ctypes.data_as(engine._i32p), grp.ctypes.data_as(engine._i32p), C.byref(cost))
        assert st == 0 and (grp >= 0).all()
        sizes = np.bincount(grp)
        assert sizes.min() >= 1 and sizes.max() <= (3 if radius.max() >= 24 else 5)   # KV_GSMALL for large-radius chains
        by_radius = [sorted(radius[grp == g]) for g in range(len(sizes))]
        assert all(a[-1] <= b[0] for a, b in zip(by_radius, by_radius[1:]))          # groups are contiguous in radius
        assert cost.value == sum(r[-1] * (2 * len(r) + 1) + len(r) for r in by_radius)
        ref_instr, exec_instr = bench.fp64_instr_per_bin(octs)
        assert exec_instr == cost.value + int(sum(3 * r + 1 for r in radius))        # axis-0 plan + axis-1 pass
        assert cost.value < sum(3 * r + 1 for r in radius)                            # sharing always pays
    assert lib.mb200_kv_plan(0, None, None, None) != 0


def test_kh_ring_plan_keeps_two_boxes_in_flight():
    """The axis-1 kernel's staging ring (host-only entry point): boxes stay inside the ring on 128-byte boundaries, a box
    only overlaps boxes of steps up to the one it declares as its dependency, and -- the point of placing large boxes at
    alternating ends of the ring -- the box of step s never has to wait for step s - 1 (its load can overlap that step)."""
    from mustache_b200 import engine, ladder
    lib = engine.load_library()
    for octs in ([1.6, 3.2], [1.6, 3.2, 6.4], [1.6, 3.2, 6.4, 12.8], [0.7, 1.4, 2.8, 5.6, 11.2], [2.3], [4.0, 9.0]):
        radius = np.array([s.radius for s in ladder.build_program(octs).steps], np.int32)
        n = len(radius)
        off, size, dep = (np.zeros(n, np.int32) for _ in range(3))
        cap = C.c_int32(0)
        ptr = lambda a: a.ctypes.data_as(engine._i32p)
        assert lib.mb200_kh_ring_plan(n, ptr(radius), ptr(off), ptr(size), ptr(dep), C.byref(cap)) == 0
        assert (off % 16 == 0).all() and (off >= 0).all() and (off + size <= cap.value).all()
        assert cap.value * 8 <= 112 * 1024
        for s in range(n):
            over = [t for t in range(s) if off[t] < off[s] + size[s] and off[s] < off[t] + size[t]]
            assert dep[s] == (max(over) if over else -1)
            if s >= 2:
                assert dep[s] <= s - 2, (octs, s, int(radius[s]))        # double buffered at the very least
    assert lib.mb200_kh_ring_plan(0, None, None, None, None, None) != 0
