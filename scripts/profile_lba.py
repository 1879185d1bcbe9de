"""Small driver for ncu: one LocalBA solve on the config-4 graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_b200 import scenes  # noqa: E402
from orb_slam3_b200.optimizer import LocalBundleAdjustment  # noqa: E402

K, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50, 20000)
g, _ = scenes.lba_graph(K, L, seed=0)
gv = scenes.lba_view(g)
lba = LocalBundleAdjustment()
for _ in range(2):
    r = lba(gv)
print(r["stats"])
