#!/usr/bin/env python3
"""Driver for ncu captures of the SURVEY.md 8(f) kernels: a few ComputeStereoMatches batches on
64 rectified 720p pairs and a few PoseOptimization batches of 128 frames x 1000 edges."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_b200 import scenes  # noqa: E402
from orb_slam3_b200.extractor import ORBextractor  # noqa: E402
from orb_slam3_b200.optimizer import PoseOptimization  # noqa: E402
from orb_slam3_b200.stereo import StereoMatcher  # noqa: E402
from orb_slam3_b200.synth import synth_frame, stereo_right  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lefts = [synth_frame(720, 1280, 9000 + i) for i in range(4)]
rights = [stereo_right(l, 9100 + i, disparities=(6, 24, 12)) for i, l in enumerate(lefts)]
el, er = ORBextractor(2000, 1.2, 8, 20, 7), ORBextractor(2000, 1.2, 8, 20, 7)
el.extract_batch([lefts[i % 4] for i in range(pairs)])
er.extract_batch([rights[i % 4] for i in range(pairs)])
sm = StereoMatcher()
for _ in range(3):
    kept, ur, dp = sm.compute_batch(el, er, pairs, 386.0, 386.0 / 700.0)
print("stereo: %.1f matches per pair, %.3f ms per %d pairs" % (kept.mean(), sm.last_ms(), pairs))
po = PoseOptimization()
views = [scenes.pose_scene(1000, seed=100 + i % 16)[0] for i in range(128)]
for _ in range(3):
    inl, pose, outs, stats = po.batch(views)
print("pose: %.1f inliers, %.1f LM trials per frame, %.3f ms per 128 frames" % (inl.mean(), stats[:, 2].mean(), po.last_ms()))
