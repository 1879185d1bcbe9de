// Which rigs take the GPU path.  The front-end kernels of liborbb200.so (matchers, PoseOptimization, isInFrustum,
// ComputeStereoMatches) implement the Pinhole projection of a single camera (fx, fy, cx, cy [, bf]); for them
// everything else -- KannalaBrandt8 (also monocular: TUM-VI), a second camera (mpCamera2), Frame::Nleft != -1 (the
// `Nleft != -1` branches of ORBmatcher.cc:144-210) -- keeps the reference bodies, which INTEGRATION.md renames to
// *_Reference.  LocalBundleAdjustment is wider: lba_solve carries both camera models and the second-camera edges
// (EdgeSE3ProjectXYZToBody, OptimizableTypes.cpp:192-213, Optimizer.cc:1366-1400), so only an unknown camera type
// falls back (lba_gpu_path).  One Atlas has one rig (System builds every Frame from the same Settings), so the
// decision is taken once per call on the object at hand, BEFORE anything is marked or erased.
#pragma once
#include "Frame.h"
#include "GeometricCamera.h"
#include "KeyFrame.h"

namespace ORB_SLAM3 {
namespace orbb200_gate {

inline bool pinhole_single(GeometricCamera* cam, GeometricCamera* cam2, int nleft) {  // GetType() is not a const member
  return cam && cam->GetType() == GeometricCamera::CAM_PINHOLE && !cam2 && nleft == -1;
}
inline bool gpu_path(const Frame& F) { return pinhole_single(F.mpCamera, F.mpCamera2, F.Nleft); }
inline bool gpu_path(const KeyFrame* kf) { return pinhole_single(kf->mpCamera, kf->mpCamera2, kf->NLeft); }
inline bool known_model(GeometricCamera* cam) {
  return cam->GetType() == GeometricCamera::CAM_PINHOLE || cam->GetType() == GeometricCamera::CAM_FISHEYE;
}
inline bool lba_gpu_path(const KeyFrame* kf) {
  return kf->mpCamera && known_model(kf->mpCamera) && (!kf->mpCamera2 || known_model(kf->mpCamera2));
}

}  // namespace orbb200_gate
}  // namespace ORB_SLAM3
