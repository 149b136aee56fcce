"""Where the end-to-end step of bench.py spends its time: upload alone, run alone, record fetch alone (config 2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mustache_b200.engine import ScaleSpaceEngine
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "2"]
eng = ScaleSpaceEngine(0); eng.set_octaves(cfg["octaves"])
tiles, keep = bench.make_host_tiles(cfg, 0)
B = cfg["blocks"]
eng.configure(cfg["n"], cfg["dpx"], B)
def t(f, n=5):
    f(); eng.sync()
    t0 = time.perf_counter()
    for _ in range(n): f(); eng.sync()
    return (time.perf_counter() - t0) / n * 1e3
def up():
    for b, x in enumerate(tiles): eng.upload_dense(b, x)
print("upload ms", t(up))
print("run ms", t(eng.run))
for pinned in (True, False):
    print("records pinned=%s ms" % pinned, t(lambda: [eng.records(b, sort=False, pinned=pinned) for b in range(B)]))
