import sys, numpy as np
sys.path.insert(0, '/root/repo')
from orb_slam3_b200.extractor import ORBextractor
from orb_slam3_b200.synth import synth_frame
from oracle import oracle
img = synth_frame(480, 640, 2)
e = ORBextractor(1000, 1.2, 8, 20, 7)
mono, k, d = e(img)
rk, rd, rm = oracle.OracleExtractor(1000).extract(img)
print("n", len(k), len(rk), "equal", mono == rm and len(k) == len(rk) and all(np.array_equal(k[f], rk[f]) for f in ("x", "y", "angle", "octave", "response")) and np.array_equal(d, rd))
