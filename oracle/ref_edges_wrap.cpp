// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's own object code for the bundle-adjustment edges and
// camera models (oracle/_ref/libref_edges.so): /root/reference/src/OptimizableTypes.cpp, src/CameraModels/Pinhole.cpp and
// src/CameraModels/KannalaBrandt8.cpp compiled UNMODIFIED, against oracle/eigencompat/ (a functional stand-in for the
// slice of Eigen / g2o these files use: the image has no Eigen) and the OpenCV / Boost / Sophus stand-ins of
// orb_slam3_b200/shim/stubs/.  What runs here is the reference's computeError(), linearizeOplus(), isDepthPositive(),
// project() and projectJac() -- its control flow and formulas as object code -- over stand-in matrix arithmetic.
// tests/test_ref_edges.py holds the oracle (orc_lba.cpp, orc_pose.cpp) against it.  Nothing in the product links this.
#include <vector>

#include "OptimizableTypes.h"   // the reference's headers (-I /root/reference/include ...)
#include "Pinhole.h"
#include "KannalaBrandt8.h"

// Pinhole / KannalaBrandt8::ReconstructWithTwoViews reference TwoViewReconstruction (src/TwoViewReconstruction.cc needs
// DBoW2's random utilities and much more of Eigen); never called from here.
namespace ORB_SLAM3 {
TwoViewReconstruction::TwoViewReconstruction(const Eigen::Matrix3f&, float, int) {}
bool TwoViewReconstruction::Reconstruct(const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&, const std::vector<int>&,
                                        Sophus::SE3f&, std::vector<cv::Point3f>&, std::vector<bool>&) { return false; }
}  // namespace ORB_SLAM3

namespace {
using namespace ORB_SLAM3;
GeometricCamera* make_camera(int model, const float* p8) {
  if (model == 1) return new KannalaBrandt8(std::vector<float>(p8, p8 + 8));
  return new Pinhole(std::vector<float>(p8, p8 + 4));
}
g2o::SE3Quat se3(const double* p) {  // quaternion x y z w + translation, as g2o::SE3Quat(q, t)
  return g2o::SE3Quat(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
template <class M> void dump(const M& J, int rows, int cols, double* out) {
  for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) out[r * cols + c] = J(r, c);
}
}  // namespace

extern "C" {

void ref_cam_project(int model, const float* p8, const double* X, double* uv) {
  GeometricCamera* cam = make_camera(model, p8);
  const Eigen::Vector2d r = cam->project(Eigen::Vector3d(X[0], X[1], X[2]));
  uv[0] = r[0]; uv[1] = r[1];
  delete cam;
}
void ref_cam_project_jac(int model, const float* p8, const double* X, double* J6) {
  GeometricCamera* cam = make_camera(model, p8);
  dump(cam->projectJac(Eigen::Vector3d(X[0], X[1], X[2])), 2, 3, J6);
  delete cam;
}

// EdgeSE3ProjectXYZ (trl == NULL) / EdgeSE3ProjectXYZToBody: err[2], Jxi = d err / d point (2x3), Jxj = d err / d pose (2x6)
void ref_edge_binary(int model, const float* p8, const double* pose7, const double* trl7, const double* X, const double* obs,
                     double* err, double* Jxi6, double* Jxj12, int* depth_pos) {
  GeometricCamera* cam = make_camera(model, p8);
  g2o::VertexSE3Expmap vpose;
  vpose.setEstimate(se3(pose7));
  g2o::VertexSBAPointXYZ vpt;
  vpt.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
  if (trl7) {
    EdgeSE3ProjectXYZToBody e;
    e.setVertex(0, &vpt); e.setVertex(1, &vpose);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.pCamera = cam; e.mTrl = se3(trl7);
    e.computeError(); e.linearizeOplus();
    err[0] = e.error()[0]; err[1] = e.error()[1];
    dump(e.jacobianOplusXi(), 2, 3, Jxi6); dump(e.jacobianOplusXj(), 2, 6, Jxj12);
    *depth_pos = e.isDepthPositive();
  } else {
    EdgeSE3ProjectXYZ e;
    e.setVertex(0, &vpt); e.setVertex(1, &vpose);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.pCamera = cam;
    e.computeError(); e.linearizeOplus();
    err[0] = e.error()[0]; err[1] = e.error()[1];
    dump(e.jacobianOplusXi(), 2, 3, Jxi6); dump(e.jacobianOplusXj(), 2, 6, Jxj12);
    *depth_pos = e.isDepthPositive();
  }
  delete cam;
}

// EdgeSE3ProjectXYZOnlyPose (trl == NULL) / EdgeSE3ProjectXYZOnlyPoseToBody (Optimizer::PoseOptimization): Jxi = d err / d pose (2x6)
void ref_edge_unary(int model, const float* p8, const double* pose7, const double* trl7, const double* Xw, const double* obs,
                    double* err, double* Jxi12, int* depth_pos) {
  GeometricCamera* cam = make_camera(model, p8);
  g2o::VertexSE3Expmap vpose;
  vpose.setEstimate(se3(pose7));
  if (trl7) {
    EdgeSE3ProjectXYZOnlyPoseToBody e;
    e.setVertex(0, &vpose);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.pCamera = cam; e.mTrl = se3(trl7); e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.computeError(); e.linearizeOplus();
    err[0] = e.error()[0]; err[1] = e.error()[1];
    dump(e.jacobianOplusXi(), 2, 6, Jxi12);
    *depth_pos = e.isDepthPositive();
  } else {
    EdgeSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &vpose);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.pCamera = cam; e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.computeError(); e.linearizeOplus();
    err[0] = e.error()[0]; err[1] = e.error()[1];
    dump(e.jacobianOplusXi(), 2, 6, Jxi12);
    *depth_pos = e.isDepthPositive();
  }
  delete cam;
}

}  // extern "C"
