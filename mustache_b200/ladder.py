"""Scale table for the scale-space engine: sigma ladder, scipy truncation radii and normalised Gaussian taps.

Host-side, numpy only.  The taps are computed here (never on the device) exactly the way the reference's
dependency builds them, because the truncation radius and the renormalisation are part of the detector
(SURVEY.md App. B):
  mustache.py:716-718, 722-724, 731-733, 748-750   sigma_k, w = 2*ceil(2 sigma)+1, t = ((w-1)/2 - 0.5)/sigma
  scipy.ndimage._filters.gaussian_filter1d          lw = int(t*sigma + 0.5)
  scipy.ndimage._filters._gaussian_kernel1d         exp(-0.5/sigma^2 * x^2) / sum
The engine walks a *chain* of Gaussians with strictly increasing sigma: L_s = g_s - g_{s+1}.  Levels 11 and 12
of octave o are bit-identical to levels 1 and 2 of octave 2o for the default ladder, so consecutive octaves
share them (22 Gaussians for 2 octaves instead of 24); when they are not bit-identical the chain is cut and
restarted, which reproduces the reference's per-octave loop exactly either way.
"""
import math
from dataclasses import dataclass, field

import numpy as np

LEVELS_PER_OCTAVE = 12  # s = 10 is hard-coded at mustache.py:711; -i/--iterations is ignored


def level_sigma(o, k, s=10):
    if k == 1:
        return o
    if k == 2:
        return o * 2 ** ((2 - 1) / s)
    if k == 3:
        return o * 2 ** ((3 - 1) / s)
    return o * 2 ** ((k - 1) / s)


def scipy_taps(sigma):
    w = 2 * math.ceil(2 * sigma) + 1
    t = (((w - 1) / 2) - 0.5) / sigma
    sd = float(sigma)
    lw = int(t * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return lw, phi


@dataclass
class Step:
    """One Gaussian of the chain.  `score_id` > 0 means: after forming L from this Gaussian, score the DoG two
    steps back ... i.e. the ring centre, and report `score_id` (= octave*12 + i, the reference's scales[o][i])."""
    sigma: float
    radius: int
    taps: np.ndarray            # full 2R+1 symmetric taps
    restart: bool               # chain is cut before this Gaussian (no DoG formed with the previous one)
    score_id: int               # 0 = do not score after this step
    score_sigma: float = 0.0    # DETECTION_SCALE reported for score_id
    diff_ref: bool = False      # this step's new DoG is the octave's L_2 (diff_mustache.py quirk: the only
                                # difference-stack DoG ever used, see diff_mustache.py:336, 371-378, 413-425)
    octave: int = 0


@dataclass
class ScaleProgram:
    steps: list = field(default_factory=list)
    sigma_of_id: dict = field(default_factory=dict)     # score_id -> sigma (python float, reported verbatim)
    octave_of_id: dict = field(default_factory=dict)

    @property
    def max_radius(self):
        return max(s.radius for s in self.steps)

    @property
    def n_scored(self):
        return sum(1 for s in self.steps if s.score_id)


def build_program(octave_values, dedupe=True):
    """Flatten the reference's `for o in octave_values: ... for i in range(3, s+2)` into a chain of steps.

    Within an octave, Gaussian k (1..12) forms L_{k-1} = G_{k-1} - G_k; once L_i exists (k = i+1, i = 3..11) the
    reference scores L_{i-1} against L_{i-2} and L_i and reports sigma_i (mustache.py:744-768).
    """
    prog = ScaleProgram()
    prev_tail = None      # (taps of level 11, taps of level 12) of the previous octave
    for oi, o in enumerate(octave_values):
        sig = [None] + [level_sigma(o, k) for k in range(1, LEVELS_PER_OCTAVE + 1)]
        taps = [None] + [scipy_taps(sg) for sg in sig[1:]]
        shared = (dedupe and prev_tail is not None
                  and all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(prev_tail, (taps[1], taps[2]))))
        first = 3 if shared else 1
        for k in range(first, LEVELS_PER_OCTAVE + 1):
            i = k - 1                   # the DoG this Gaussian completes is L_i
            sid = oi * LEVELS_PER_OCTAVE + i if 3 <= i <= 11 else 0
            st = Step(sigma=sig[k], radius=taps[k][0], taps=taps[k][1], restart=(k == 1), score_id=sid,
                      score_sigma=sig[i] if sid else 0.0, diff_ref=(k == 3), octave=oi)
            prog.steps.append(st)
            if sid:
                prog.sigma_of_id[sid] = sig[i]
                prog.octave_of_id[sid] = oi
        prev_tail = (taps[11], taps[12])
    return prog


def build_diff_program(octave_values):
    """Chain for the difference stack of diff_mustache: per octave only G_2 and G_3 are ever used, because the
    reference computes `Lc = Gc - Gn` once before the level loop (diff_mustache.py:336) and never rotates it
    (diff_mustache.py:413-425), so norm.fit / norm.cdf (:371-378) always see that octave's L_2."""
    prog = ScaleProgram()
    for oi, o in enumerate(octave_values):
        for k in (2, 3):
            sg = level_sigma(o, k)
            r, taps = scipy_taps(sg)
            prog.steps.append(Step(sigma=sg, radius=r, taps=taps, restart=(k == 2), score_id=0, diff_ref=(k == 3), octave=oi))
    return prog
