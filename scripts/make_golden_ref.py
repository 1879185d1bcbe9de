#!/usr/bin/env python3
"""Generate tests/golden/ref_*.npz from the REFERENCE's own object code (oracle/_ref: /root/reference/src/
ORBextractor.cc compiled unmodified, `make -C oracle ref`).  Run in the build container, where /root/reference
exists.  Unlike the other fixtures (oracle outputs, scripts/make_golden.py) these are reference vectors:

  ref_extract_640x480.npz  : BASELINE.json configs[0] -- ORBextractor(1000, 1.2, 8, 20, 7) on synth_frame(480, 640, 1):
                             keypoints (all seven KeyPoint fields, output order), 32-byte descriptors, monoIndex
  ref_extract_lowtex.npz   : the same extractor on a low-texture frame (minThFAST fallback cells) with a lapping
                             area [200, 420] (stereo-fisheye ordering of operator())
  ref_extract_1280x720.npz : configs[1] frame synth_frame(720, 1280, 3), 2000 features -- stored as a SHA-256 of the
                             keypoint / descriptor bytes + counts (the arrays themselves are 120 KB)
  ref_match_helpers.npz    : ORBmatcher::DescriptorDistance on 512 random pairs, ComputeThreeMaxima on 64 histograms
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from orb_slam3_b200.synth import synth_frame  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
assert ref.build() is not None, "needs /root/reference"

k, d, mono = ref.RefExtractor(1000).extract(synth_frame(480, 640, 1))
np.savez_compressed(os.path.join(out, "ref_extract_640x480.npz"), kps=k, desc=d, mono=mono)
k, d, mono = ref.RefExtractor(1000).extract(synth_frame(480, 640, 11, low_texture=True), lap=(200, 420))
np.savez_compressed(os.path.join(out, "ref_extract_lowtex.npz"), kps=k, desc=d, mono=mono, lap=np.array([200, 420]))
k, d, mono = ref.RefExtractor(2000).extract(synth_frame(720, 1280, 3))
np.savez_compressed(os.path.join(out, "ref_extract_1280x720.npz"), n=len(k), mono=mono,
                    sha_kps=np.array(hashlib.sha256(k.tobytes()).hexdigest()),
                    sha_desc=np.array(hashlib.sha256(d.tobytes()).hexdigest()))
rng = np.random.default_rng(5)
a = rng.integers(0, 256, (512, 32), dtype=np.uint8)
b = rng.integers(0, 256, (512, 32), dtype=np.uint8)
b[:16] = a[:16]
b[16:32] = ~a[16:32]
dist = np.array([ref.descriptor_distance(x, y) for x, y in zip(a, b)], np.int32)
hist = rng.integers(0, 40, (64, 30)).astype(np.int32)
hist[:8] = (hist[:8] > 36) * hist[:8]          # sparse histograms: the 10 % rules fire
hist[8] = 0
tm = np.array([ref.three_maxima(h) for h in hist], np.int32)
np.savez_compressed(os.path.join(out, "ref_match_helpers.npz"), a=a, b=b, dist=dist, hist=hist, three=tm)
print("wrote reference vectors to", out)
