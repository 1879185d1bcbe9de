"""Small driver for ncu / stage timing: extract a batch of synthetic 720p frames a few times.
usage: profile_extract.py [batch] [reps] [--device]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_frames, H, W  # noqa: E402
from orb_slam3_b200.extractor import ORBextractor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames, _ = make_frames(min(B, 17), 1)
frames = np.stack([frames[i % len(frames)] for i in range(B)])
ext = ORBextractor(2000, 1.2, 8, 20, 7)
if "--device" in sys.argv:
    import torch
    d = torch.from_numpy(frames).cuda()
    ext.extract_batch_device(d.data_ptr(), B, H, W, W, H * W)
    ext.synchronize()
    ext.set_profiling(True)
    ext.stage_times(reset=True)
    for _ in range(reps):
        ext.extract_batch_device(d.data_ptr(), B, H, W, W, H * W)
    ext.synchronize()
    st = ext.stage_times()
    print({k: round(v[0] / reps * 1e3 / B, 2) for k, v in st.items()}, "us/frame, batch", B)
else:
    for _ in range(reps):
        res = ext.extract_batch(list(frames))
    print("keypoints", sum(len(r[1]) for r in res))
