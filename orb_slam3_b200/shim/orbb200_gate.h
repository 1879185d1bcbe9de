// Which rigs take the GPU path.  The kernels of liborbb200.so implement the Pinhole projection of a single
// camera (fx, fy, cx, cy [, bf]); everything else -- KannalaBrandt8 (also monocular: TUM-VI), a second camera
// (mpCamera2: EdgeSE3ProjectXYZToBody right-camera edges, OptimizableTypes.cpp:192-213, Optimizer.cc:1366-1400),
// Frame::Nleft != -1 (the `Nleft != -1` branches of ORBmatcher.cc:144-210) -- keeps the reference bodies, which
// INTEGRATION.md renames to *_Reference.  One Atlas has one rig (System builds every Frame from the same
// Settings), so the decision is taken once per call on the object at hand, BEFORE anything is marked or erased.
#pragma once
#include "Frame.h"
#include "GeometricCamera.h"
#include "KeyFrame.h"

namespace ORB_SLAM3 {
namespace orbb200_gate {

inline bool pinhole_single(const GeometricCamera* cam, const GeometricCamera* cam2, int nleft) {
  return cam && cam->GetType() == GeometricCamera::CAM_PINHOLE && !cam2 && nleft == -1;
}
inline bool gpu_path(const Frame& F) { return pinhole_single(F.mpCamera, F.mpCamera2, F.Nleft); }
inline bool gpu_path(const KeyFrame* kf) { return pinhole_single(kf->mpCamera, kf->mpCamera2, kf->NLeft); }

}  // namespace orbb200_gate
}  // namespace ORB_SLAM3
