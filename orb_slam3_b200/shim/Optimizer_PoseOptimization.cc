// Replacement for Optimizer::PoseOptimization (reference src/Optimizer.cc:814-1115, SURVEY.md
// 8(f-2)).  The edge set-up loop under MapPoint::mGlobalMutex and the write-back are the
// reference's own logic on its own data structures; the four rounds of optimizer.optimize(10)
// with their chi2 re-classification are pose_optimize() of liborbb200.so (one kernel launch).
// Build inside the ORB_SLAM3 tree in place of that one function (the reference body goes under
// `#ifndef ORB_B200_HOTPATH`, kept as PoseOptimization_Reference for the fisheye-stereo rig).
// Syntax-checked against the reference's headers over stand-ins for its third-party libraries (tests/test_shim_syntax.py); not linked here -- see INTEGRATION.md.
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "Optimizer.h"
#include "orb_b200.h"
#include "orbb200_gate.h"

namespace ORB_SLAM3 {

int Optimizer::PoseOptimization(Frame* pFrame) {
  if (!orbb200_gate::gpu_path(*pFrame)) return PoseOptimization_Reference(pFrame);  // KB8 / rigid-body stereo: pCamera->project edges
  const int N = pFrame->N;
  std::vector<float> xw, obs, inv_sigma2;
  std::vector<size_t> index;  // edge -> keypoint i (vnIndexEdgeMono / vnIndexEdgeStereo merged, keypoint order)
  xw.reserve(3 * N); obs.reserve(3 * N); inv_sigma2.reserve(N); index.reserve(N);
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);  // :856
    for (int i = 0; i < N; i++) {
      MapPoint* pMP = pFrame->mvpMapPoints[i];
      if (!pMP) continue;
      pFrame->mvbOutlier[i] = false;                            // :869, :903
      const cv::KeyPoint& kpUn = pFrame->mvKeysUn[i];
      const Eigen::Vector3f X = pMP->GetWorldPos();
      xw.insert(xw.end(), {X.x(), X.y(), X.z()});
      obs.insert(obs.end(), {kpUn.pt.x, kpUn.pt.y, pFrame->mvuRight[i]});  // < 0: monocular edge (:867)
      inv_sigma2.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
      index.push_back(i);
    }
  }
  const int nInitialCorrespondences = (int)index.size();
  if (nInitialCorrespondences < 3) return 0;                    // :1000-1001

  pose_opt_view v;
  v.n = nInitialCorrespondences;
  v.xw = xw.data(); v.obs = obs.data(); v.inv_sigma2 = inv_sigma2.data();
  v.fx = pFrame->fx; v.fy = pFrame->fy; v.cx = pFrame->cx; v.cy = pFrame->cy; v.bf = pFrame->mbf;
  const Sophus::SE3<float> Tcw = pFrame->GetPose();             // :829-831
  const Eigen::Quaterniond q = Tcw.unit_quaternion().cast<double>();
  const Eigen::Vector3d t = Tcw.translation().cast<double>();
  v.pose[0] = q.x(); v.pose[1] = q.y(); v.pose[2] = q.z(); v.pose[3] = q.w();
  v.pose[4] = t.x(); v.pose[5] = t.y(); v.pose[6] = t.z();

  static thread_local orb_poseopt* h = nullptr;                 // Tracking thread only
  if (!h && poseopt_create(/*device=*/0, &h) != ORB_OK)
    throw std::runtime_error(std::string("poseopt_create: ") + orb_last_error());
  double pose[7];
  std::vector<uint8_t> outlier(v.n);
  const int inliers = pose_optimize(h, &v, pose, outlier.data());
  if (inliers < 0) throw std::runtime_error(std::string("pose_optimize: ") + orb_last_error());

  for (int e = 0; e < v.n; e++) pFrame->mvbOutlier[index[e]] = outlier[e] != 0;
  // :1106-1112
  const Eigen::Quaterniond qr(pose[3], pose[0], pose[1], pose[2]);
  const Eigen::Vector3d tr(pose[4], pose[5], pose[6]);
  pFrame->SetPose(Sophus::SE3<float>(qr.cast<float>(), tr.cast<float>()));
  return inliers;  // nInitialCorrespondences - nBad
}

}  // namespace ORB_SLAM3
