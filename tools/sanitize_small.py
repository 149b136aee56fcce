#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the library on tiny inputs --
single map (three-kernel and fused path), differential, BH + candidate selection, normaliser (both branches), batch COO
upload.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mustache_b200 import blockrun, normalize, synth as gen  # noqa: E402
from mustache_b200.engine import ScaleSpaceEngine  # noqa: E402


def main():
    eng = ScaleSpaceEngine(0)
    n, dpx = 300, 120
    tiles = [gen.band_to_dense(gen.dense_band_tile(n, dpx, seed=5 + b, blob_seed=9 + b, nblobs=6, missing=0.1), n) for b in range(2)]
    masks = []
    for t in tiles:
        r, c = np.nonzero(np.triu(t, 4))
        masks.append((r, c, t[r, c]))
    for octs in ([1.6, 3.2], [1.6, 3.2, 6.4, 12.8]):
        for fused in (False, True):
            eng.set_octaves(octs)
            eng.set_fusion(fused)
            eng.configure(n, dpx, 2)
            eng.upload_coo_batch(0, *blockrun.concat_coo(masks))
            eng.run()
            eng.select_candidates(0.3, 0.5)
            c = eng.candidates_batch()
            print("octaves", len(octs), "fused", fused, "records", [r["n_found"] for r in eng.records_batch()], "candidates", [len(x["rows"]) for x in c])
    eng.set_fusion(False)
    eng.set_octaves([1.6, 3.2], differential=True)
    eng.configure(n, dpx, 2)
    eng.upload_dense(0, tiles[0])
    eng.upload_dense(1, tiles[1])
    eng.run_differential()
    eng.select_candidates(0.3, 0.5)
    print("differential candidates", [len(x["rows"]) for x in eng.candidates_batch(pair=True)])
    g, l = eng.debug_level(0, 5)
    print("debug level", float(np.abs(g).max()) > 0)
    for res, dist in ((5000, 400), (50000, 40)):
        x, y, cnt = gen.synthetic_chromosome(900, dist, 18.0, seed=3, nloops=5, loop_dmax=20)
        v = cnt.astype(np.float64) * 1.37
        normalize.normalize_sparse_device(eng, x, y, v, res, dist)
        print("normaliser", res, float(np.abs(v).max()))
    eng.close()
    print("sanitize run ok")


if __name__ == "__main__":
    main()
