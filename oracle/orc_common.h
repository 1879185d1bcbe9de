// TEST INFRASTRUCTURE ONLY -- CPU oracle for the orb_slam3_b200 hot path.
//
// Nothing in the product (orb_slam3_b200/, include/) may include, link or call
// anything in this directory.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker and
// as the timed CPU baseline.
//
// PARITY UNPINNED BY THE REFERENCE: ORB_SLAM3 ships no tests, golden vectors or
// known-answer fixtures for this path and the
// reference itself cannot be compiled in this image (OpenCV C++, Eigen,
// Pangolin and Boost headers are absent).  The oracle is therefore pinned to
//   * cv2 4.13 (Python) for the un-vendored OpenCV arithmetic (resize, FAST,
//     GaussianBlur, fastAtan2) -- tests/test_oracle_vs_cv2.py + tests/golden/
//   * the reference sources it restates, cited per function as file:line
//     relative to /root/reference.
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
