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
//   * matcher rows (a10, a11, a13), Frame::isInFrustum, Frame::ComputeStereoMatches: the reference's own object code --
//     ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc compiled UNMODIFIED against cvcompat/ + eigencompat/ into
//     oracle/_ref/libref_front.so, real Frame / KeyFrame / MapPoint objects -- tests/test_ref_front.py (exact);
//   * BA edges (EdgeSE3ProjectXYZ / ToBody / OnlyPose) and both camera models: OptimizableTypes.cpp, Pinhole.cpp,
//     KannalaBrandt8.cpp as object code (libref_edges.so) -- tests/test_ref_edges.py (1e-9);
//   * Frame::ComputeBoW's transform: the vendored DBoW2 as object code (libref_bow.so) -- tests/test_ref_bow.py (exact);
//   * g2o's stereo edges (EdgeStereoSE3ProjectXYZ, ...OnlyPose), SE3Quat (exp / oplus) and RobustKernelHuber: the vendored
//     types_six_dof_expmap.{h,cpp}, se3quat.h, robust_kernel*.cpp as object code (libref_g2o.so) -- tests/test_ref_edges.py;
//   * the LM control law of all three optimisers (LocalBundleAdjustment, PoseOptimization, LocalInertialBA):
//     optimization_algorithm_levenberg.cpp as object code (libref_lm.so) driving this oracle's own operations through
//     orc_lm_ops.h -- tests/test_ref_lm.py (bit for bit);
//   * the accumulation of the edges into the normal equations: constructQuadraticForm() of base_binary_edge.hpp /
//     base_unary_edge.hpp as object code (libref_g2o.so) over whole windows / frames -- tests/test_ref_edges.py (1e-12);
//   * the visual edges of LocalInertialBA (EdgeMono / EdgeStereo): src/G2oTypes.cc as object code (libref_vi.so) --
//     tests/test_ref_edges.py (1e-12);
//   * g2o's block solver (Schur complement, sparse LDL^T) and the inertial edges of G2oTypes.cc: "parity unpinned" by the
//     reference (those translation units need the real Eigen); pinned by independent numpy restatements, finite
//     differences and dense solves, with the reference lines cited per function.
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
