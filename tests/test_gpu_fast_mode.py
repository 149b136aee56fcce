"""Opt-in FMA fast mode (mb200_set_arithmetic(1), MUSTACHE_FAST=1): one fused multiply-add per tap instead of scipy's
multiply-then-add.  It is NOT the reference's arithmetic (Gaussians differ in the last bits), so it is gated here on
BASELINE.json's own tolerance -- identical bin coordinates and detection scales, FDR within 1e-6 -- on every golden input
that runs through the CLI, and on how far a Gaussian can move."""
import os

import numpy as np
import pytest

from mustache_b200 import mustache as mm
from mustache_b200 import synth as gen
from tests import synth
from tests.test_gpu_e2e import G, _read_tsv, _same_tsv
from tests.test_gpu_configs import _cli_rows, _same_rows

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fast_engine():
    eng = mm.get_engine()
    eng.set_arithmetic(True)
    yield eng
    eng.set_arithmetic(False)


def test_fast_gaussians_stay_within_rounding(fast_engine):
    eng = fast_engine
    n, dpx, octs = 320, 140, [1.6, 3.2, 6.4, 12.8]
    c = synth.make_tile(**synth.SYNTH_TILES["n320_o4"]["gen"])
    eng.set_octaves(octs)
    mm._PROGRAM_KEY[id(eng)] = tuple(octs)
    eng.configure(n, dpx, 1)
    eng.upload_dense(0, c)
    worst = 0.0
    for step in (0, 7, 20, len(eng.program.steps) - 1):
        eng.set_arithmetic(True)
        gf, _ = eng.debug_level(0, step)
        eng.set_arithmetic(False)
        ge, _ = eng.debug_level(0, step)
        live = ge != 0
        assert live.any() and not np.array_equal(gf, ge)         # it really is other arithmetic ...
        worst = max(worst, float(np.abs(gf[live] - ge[live]).max() / np.abs(ge[live]).max()))
    assert worst < 1e-14                                          # ... a few ulps away
    eng.set_arithmetic(True)


@pytest.mark.parametrize("name", list(synth.SYNTH_TILES))
def test_fast_mustache_function(fast_engine, name):
    spec = synth.SYNTH_TILES[name]
    z = np.load(os.path.join(G, "synth_%s.npz" % name))
    c = synth.make_tile(**spec["gen"])
    loops = mm.mustache(c, "1", "1", 5000, [], 0, c.shape[0], -1, spec["dpx"], list(spec["octaves"]), spec["st"], spec["pt"])
    got, ref = np.array(loops, float).reshape(-1, 4), z["loops"]
    assert got.shape == ref.shape and np.array_equal(got[:, [0, 1, 3]], ref[:, [0, 1, 3]])
    assert np.abs(got[:, 2] - ref[:, 2]).max() <= 1e-6


def test_fast_cli_chr21(fast_engine, tmp_path):
    raw, kr = synth.write_chr21_text(str(tmp_path))
    out = str(tmp_path / "chr21_fast.tsv")
    mm.main(["-f", raw, "-b", kr, "-ch", "21", "-r", "5kb", "-pt", "0.1", "-st", "0.8", "-o", out])
    assert _same_tsv(out, os.path.join(G, "chr21_loops.tsv")) == 90


def test_fast_cli_config4_and_dense_1kb(fast_engine, tmp_path):
    names = ["s1", "s2", "s3"]
    key = lambda r: (r[0], int(r[1]), int(r[4]))
    ref = sorted([r for r in _read_tsv(os.path.join(G, "cfg4_loops.tsv")) if r[0] in names], key=key)
    (tmp_path / "a").mkdir()
    got = _cli_rows(tmp_path / "a", names, 5000, [(k, gen.CONFIG4[k]) for k in names])
    _same_rows(sorted(got, key=key), ref)
    (tmp_path / "b").mkdir()
    _same_rows(_cli_rows(tmp_path / "b", ["chrT"], 1000, [("chrT", gen.CONFIG3D)]), _read_tsv(os.path.join(G, "cfg3d_loops.tsv")))
