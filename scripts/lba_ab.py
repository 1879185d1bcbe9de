#!/usr/bin/env python3
"""A/B of the reduced-solve kernels on the GPU: ORB_B200_LDLT=dense|sky per subprocess (the switch is read once).
Usage: python scripts/lba_ab.py [K L]...   -> one JSON line per (config, solver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, sys, numpy as np
sys.path.insert(0, %r)
from orb_slam3_b200 import scenes
from orb_slam3_b200.optimizer import LocalBundleAdjustment
K, L, shuffle = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g, _ = scenes.lba_graph(K, L, seed=0)
if shuffle:
    g = scenes.permute_keyframes(g, np.random.default_rng(1).permutation(len(g["kf_fixed"])))
lba = LocalBundleAdjustment()
gv = scenes.lba_view(g)
best = None
for rep in range(4):
    st = lba(gv)["stats"]
    if rep and (best is None or st["ms_total"] < best["ms_total"]):
        best = st
print(json.dumps({k: best[k] for k in ("iterations", "trials", "ms_total", "ms_linearize", "ms_schur", "ms_solve", "ms_update",
                                       "solver_kind", "envelope_rows_max", "ms_host_prep", "ms_wall", "chi2_final", "n_pairs")}))
''' % ROOT

cfgs = [(50, 20000), (200, 80000)]
if len(sys.argv) > 2:
    cfgs = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
for K, L in cfgs:
    for mode, shuffle in (("dense", 0), ("sky", 0), ("win", 0), ("auto", 1)):
        env = dict(os.environ)
        if mode != "auto":
            env["ORB_B200_LDLT"] = mode
        r = subprocess.run([sys.executable, "-c", CHILD, str(K), str(L), str(shuffle)], capture_output=True, text=True, env=env)
        print(json.dumps({"K": K, "L": L, "mode": mode, "shuffled_keyframes": shuffle,
                          "result": json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-600:]}), flush=True)
