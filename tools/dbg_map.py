import numpy as np, sys, ctypes as C
sys.path.insert(0,".")
from mustache_b200.engine import ScaleSpaceEngine
e=ScaleSpaceEngine(0); e.set_octaves([1.6,3.2]); e.configure(256,100,1)
buf=(C.c_uint64*16)()
e.lib.mb200_debug_tensormap.restype=C.c_int; e.lib.mb200_debug_tensormap.argtypes=[C.c_void_p,C.c_int,C.c_void_p]
for w in (0,-1):
    e.lib.mb200_debug_tensormap(e.h, w, buf)
    print("engine map", w, " ".join("%016x"%x for x in buf))
