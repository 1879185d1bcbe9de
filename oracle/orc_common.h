// TEST INFRASTRUCTURE ONLY -- CPU oracle for the orb_slam3_b200 hot path.
//
// Nothing in the product (orb_slam3_b200/, include/) may include, link or call
// anything in this directory.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker and
// as the timed CPU baseline.
//
// PARITY PINNING.  ORB_SLAM3 ships no tests, golden vectors or known-answer fixtures for this path
// (SURVEY.md section 4, 8c), and the whole reference cannot be built in this image (OpenCV C++, Eigen,
// Pangolin and Boost headers are absent).  What pins the oracle:
//   * extractor rows (a1-a9): the REFERENCE's own object code -- /root/reference/src/ORBextractor.cc compiled
//     UNMODIFIED against oracle/cvcompat/ (`make ref` -> oracle/_ref/), plus ORBmatcher::DescriptorDistance /
//     ComputeThreeMaxima -- compared with this restatement by tests/test_ref_parity.py, and reference vectors
//     generated from it under tests/golden/ref_*.npz (scripts/make_golden_ref.py);
//   * the un-vendored OpenCV arithmetic underneath (resize, FAST, GaussianBlur, fastAtan2): cv2 4.13 (Python),
//     bit-exact -- tests/test_oracle_vs_cv2.py + tests/golden/primitives_cv2.npz;
//   * matchers / g2o / DBoW2 rows: "parity unpinned" by the reference (their translation units need Eigen /
//     Sophus); pinned by independent numpy restatements and the reference sources cited per function.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Same 28-byte layout as cv::KeyPoint.
typedef struct orc_keypoint {
  float x, y;
  float size;
  float angle;
  float response;
  int32_t octave;
  int32_t class_id;
} orc_keypoint;

#ifdef __cplusplus
}
#endif
