"""Pins the CPU oracle (oracle/) against the installed scipy and against dumps of the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import postprocess as opost
from oracle import scalespace as osc
from tests import synth

G = synth.GOLDEN


def test_taps_match_scipy_kernel():
    from scipy.ndimage import _filters
    for lv in osc.sigma_ladder([1.6, 3.2, 6.4, 12.8]):
        lw = int(lv["truncate"] * float(lv["sigma"]) + 0.5)
        ref = _filters._gaussian_kernel1d(lv["sigma"], 0, lw)[::-1]
        assert lv["radius"] == lw
        assert np.array_equal(lv["taps"], ref)
        assert np.array_equal(lv["taps"], lv["taps"][::-1])          # exactly symmetric


def test_radii_table_survey_appendix_b():
    radii = [lv["radius"] for lv in osc.sigma_ladder([1.6, 3.2, 6.4, 12.8])]
    assert radii[:12] == [4, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 7]
    assert radii[12:24] == [7, 7, 8, 8, 9, 10, 10, 11, 12, 12, 13, 14]
    assert radii[24:36] == [13, 14, 15, 16, 17, 19, 20, 21, 23, 24, 26, 28]
    assert radii[36:48] == [26, 28, 29, 32, 34, 36, 39, 42, 45, 48, 52, 55]


@pytest.mark.parametrize("shape", [(97, 113), (256, 256)])
def test_gaussian_bit_exact_vs_scipy(shape):
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(5)
    a = rng.standard_normal(shape)
    for lv in osc.sigma_ladder([1.6, 3.2, 6.4])[::3]:
        if 2 * lv["radius"] >= min(shape):
            continue
        ref = gaussian_filter(a, lv["sigma"], truncate=lv["truncate"], order=0)
        got = osc.gaussian_level(a, lv["taps"])
        assert np.array_equal(ref, got), lv["sigma"]


def test_max3x3_vs_scipy():
    from scipy.ndimage import maximum_filter
    rng = np.random.default_rng(6)
    a = rng.standard_normal((64, 80)) - 0.5
    assert np.array_equal(osc.max3x3_zero(a), maximum_filter(a, footprint=np.ones((3, 3)), mode="constant"))


def test_expon_sf_matches_scipy():
    from scipy.stats import expon
    rng = np.random.default_rng(7)
    x = np.abs(rng.standard_normal(1000))
    loc, sc = expon.fit(x)
    assert loc == x.min() and sc == x.mean() - x.min()
    assert np.array_equal(1 - expon.cdf(x, loc, sc), osc.expon_sf(x, loc, sc))


def _check_against_dump(res, z, prefix=""):
    found = res["p"] != 2
    assert res["nz_count"] == int(z[prefix + "nz_count"])
    assert np.array_equal(res["rows"][found], z[prefix + "rows"])
    assert np.array_equal(res["cols"][found], z[prefix + "cols"])
    assert np.array_equal(res["scale"][found], z[prefix + "scale"])
    assert np.array_equal(res["v"][found], z[prefix + "v"])
    assert np.array_equal(res["p"][found], z[prefix + "p"])


@pytest.mark.parametrize("name", list(synth.SYNTH_TILES))
@pytest.mark.parametrize("use_scipy", [False, True])
def test_scale_space_matches_reference_dump(name, use_scipy):
    spec = synth.SYNTH_TILES[name]
    z = np.load(os.path.join(G, "synth_%s.npz" % name))
    c = synth.make_tile(**spec["gen"])
    import hashlib
    assert np.array_equal(np.frombuffer(hashlib.sha256(c.tobytes()).digest(), np.uint8), z["tile_digest"])
    res = osc.scale_space(c, spec["dpx"], spec["octaves"], use_scipy=use_scipy)
    _check_against_dump(res, z)
    # and the dense post-processing restatement reproduces the loops mustache() returned
    nz, filled = osc.mask_and_fill(c, spec["dpx"])
    loops = opost.loops_dense(filled, nz, res["p"], res["scale"], 0, spec["dpx"], spec["st"], spec["pt"])
    ref = z["loops"]
    assert len(loops) == len(ref)
    assert np.array_equal(np.array(loops, float).reshape(-1, 4), ref)


def test_diff_scale_space_matches_reference_dump():
    spec = synth.SYNTH_DIFF
    z = np.load(os.path.join(G, "diff_synth.npz"))
    c1, c2 = synth.make_pair(**spec["gen"])
    res = osc.scale_space_diff(c1, c2, spec["dpx"], spec["octaves"], use_scipy=True)
    for key, pre in (("map1", "m1_"), ("map2", "m2_")):
        st = res[key]
        found = st["p"] != 2
        assert np.array_equal(st["rows"][found], z[pre + "rows"])
        assert np.array_equal(st["cols"][found], z[pre + "cols"])
        assert np.array_equal(st["v"][found], z[pre + "v"])
        assert np.array_equal(st["scale"][found], z[pre + "scale"])
        assert np.array_equal(st["p"][found], z[pre + "p"])
        assert np.array_equal(st["pair"][found], z[pre + "pair"])


def test_bh_forms_agree():
    rng = np.random.default_rng(8)
    p = rng.random(5000) ** 3
    p[100:110] = p[100]
    a, b = opost.bh(p), opost.bh_statsmodels_form(p)
    assert np.allclose(a, b, rtol=1e-14, atol=0)
    assert (b <= 1).all() and (b >= p).all()
