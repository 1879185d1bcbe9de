"""Small driver for ncu: extract a batch of synthetic 720p frames a few times."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_frames, H, W  # noqa: E402
from orb_slam3_b200.extractor import ORBextractor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames, _ = make_frames(B, 1)
ext = ORBextractor(2000, 1.2, 8, 20, 7)
for _ in range(reps):
    res = ext.extract_batch(list(frames))
print("keypoints", sum(len(r[1]) for r in res))
