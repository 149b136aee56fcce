import numpy as np, sys
sys.path.insert(0,".")
from mustache_b200.engine import ScaleSpaceEngine
from mustache_b200 import synth as gen
e=ScaleSpaceEngine(0); e.set_octaves([1.6,3.2])
c=gen.band_to_dense(gen.dense_band_tile(256,100,seed=1,blob_seed=2,nblobs=4),256)
e.configure(256,100,1); e.upload_dense(0,c)
try:
    e.run(); e.sync(); print("run ok", e.counts(0))
except Exception as ex: print("ERR", ex)
