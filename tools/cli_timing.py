"""Wall time of the README chr21 command through the product CLI (host and device normaliser), for DESIGN.md."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import synth
from mustache_b200 import mustache as mm
d = tempfile.mkdtemp()
raw, kr = synth.write_chr21_text(d)
for mode in ("host", "device", "host", "device"):
    os.environ["MUSTACHE_NORMALIZE"] = mode
    t0 = time.perf_counter()
    mm.main(["-f", raw, "-b", kr, "-ch", "21", "-r", "5kb", "-pt", "0.1", "-st", "0.8", "-o", os.path.join(d, "o.tsv"), "-v", ""])
    print("CLI chr21 normaliser=%s wall %.2f s" % (mode, time.perf_counter() - t0))
