"""Wall time of the README chr21 command through the product CLI, piece by piece, for DESIGN.md:
reader (native / pandas) x normaliser (device / numpy)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import synth
from mustache_b200 import mustache as mm
d = tempfile.mkdtemp()
raw, kr = synth.write_chr21_text(d)
for reader, norm in (("native", "device"), ("pandas", "host"), ("native", "device"), ("pandas", "host"), ("native", "host")):
    os.environ["MUSTACHE_READER"] = reader
    os.environ["MUSTACHE_NORMALIZE"] = norm
    t0 = time.perf_counter()
    mm.main(["-f", raw, "-b", kr, "-ch", "21", "-r", "5kb", "-pt", "0.1", "-st", "0.8", "-o", os.path.join(d, "o.tsv"), "-v", ""])
    print("CLI chr21 reader=%s normaliser=%s wall %.2f s" % (reader, norm, time.perf_counter() - t0))
