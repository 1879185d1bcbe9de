// TEST INFRASTRUCTURE ONLY -- C entry points over the object code of the reference's VENDORED g2o edge types
// (oracle/_ref/libref_g2o.so): /root/reference/Thirdparty/g2o/g2o/types/types_six_dof_expmap.h and .cpp, piped UNMODIFIED
// into the compiler (oracle/Makefile) with oracle/eigencompat/g2o_unit/ standing in for the g2o core headers they include
// by relative path (functional SE3Quat / vertex / edge bases of eigencompat/g2o_base.h; the image has no Eigen).  This file
// is appended to the same header in a second pipe, so it sees the reference's own class declarations.
//   g2o::EdgeStereoSE3ProjectXYZ          (LocalBundleAdjustment's stereo edge, Optimizer.cc:1338-1362; row a16)
//   g2o::EdgeStereoSE3ProjectXYZOnlyPose  (PoseOptimization's stereo edge, Optimizer.cc:897-935; row 8f-2)
//   g2o::EdgeSE3ProjectXYZ / OnlyPose     (g2o's own pinhole edges, the formulas OptimizableTypes.cpp generalises)
// computeError(), linearizeOplus(), isDepthPositive(), chi2() run as the reference's object code.
// tests/test_ref_edges.py holds the oracle (orc_lba.cpp, orc_pose.cpp) against them.  Nothing in the product links this.
namespace {
g2o::SE3Quat se3(const double* p) {  // quaternion x y z w + translation
  return g2o::SE3Quat(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
template <class M> void dump(const M& J, int rows, int cols, double* out) {
  for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) out[r * cols + c] = J(r, c);
}
}  // namespace

extern "C" {

// stereo != 0: EdgeStereoSE3ProjectXYZ, obs[3], err[3], Jxi 3x3, Jxj 3x6; else EdgeSE3ProjectXYZ (2 rows).  info = the
// scalar the information matrix is the identity times (invSigma2); chi2_out = e' info e
void ref_g2o_edge_binary(int stereo, const double* k5, const double* pose7, const double* X, const double* obs, double info,
                         double* err, double* Jxi, double* Jxj, int* depth_pos, double* chi2_out) {
  g2o::VertexSE3Expmap vpose;
  vpose.setEstimate(se3(pose7));
  g2o::VertexSBAPointXYZ vpt;
  vpt.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
  if (stereo) {
    g2o::EdgeStereoSE3ProjectXYZ e;
    e.setVertex(0, &vpt); e.setVertex(1, &vpose);
    e.setMeasurement(Eigen::Vector3d(obs[0], obs[1], obs[2]));
    e.setInformation(Eigen::Matrix3d::Identity() * info);
    e.fx = k5[0]; e.fy = k5[1]; e.cx = k5[2]; e.cy = k5[3]; e.bf = k5[4];
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 3; i++) err[i] = e.error()[i];
    dump(e.jacobianOplusXi(), 3, 3, Jxi); dump(e.jacobianOplusXj(), 3, 6, Jxj);
    *depth_pos = e.isDepthPositive(); *chi2_out = e.chi2();
  } else {
    g2o::EdgeSE3ProjectXYZ e;
    e.setVertex(0, &vpt); e.setVertex(1, &vpose);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.setInformation(Eigen::Matrix2d::Identity() * info);
    e.fx = k5[0]; e.fy = k5[1]; e.cx = k5[2]; e.cy = k5[3];
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 2; i++) err[i] = e.error()[i];
    dump(e.jacobianOplusXi(), 2, 3, Jxi); dump(e.jacobianOplusXj(), 2, 6, Jxj);
    *depth_pos = e.isDepthPositive(); *chi2_out = e.chi2();
  }
}

// stereo != 0: EdgeStereoSE3ProjectXYZOnlyPose (Jxi 3x6), else EdgeSE3ProjectXYZOnlyPose (2x6)
void ref_g2o_edge_unary(int stereo, const double* k5, const double* pose7, const double* Xw, const double* obs, double info,
                        double* err, double* Jxi, int* depth_pos, double* chi2_out) {
  g2o::VertexSE3Expmap vpose;
  vpose.setEstimate(se3(pose7));
  if (stereo) {
    g2o::EdgeStereoSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &vpose);
    e.setMeasurement(Eigen::Vector3d(obs[0], obs[1], obs[2]));
    e.setInformation(Eigen::Matrix3d::Identity() * info);
    e.fx = k5[0]; e.fy = k5[1]; e.cx = k5[2]; e.cy = k5[3]; e.bf = k5[4];
    e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 3; i++) err[i] = e.error()[i];
    dump(e.jacobianOplusXi(), 3, 6, Jxi);
    *depth_pos = e.isDepthPositive(); *chi2_out = e.chi2();
  } else {
    g2o::EdgeSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &vpose);
    e.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
    e.setInformation(Eigen::Matrix2d::Identity() * info);
    e.fx = k5[0]; e.fy = k5[1]; e.cx = k5[2]; e.cy = k5[3];
    e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 2; i++) err[i] = e.error()[i];
    dump(e.jacobianOplusXi(), 2, 6, Jxi);
    *depth_pos = e.isDepthPositive(); *chi2_out = e.chi2();
  }
}

// VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:64-67): estimate <- SE3Quat::exp(update) * estimate, with the reference's
// own se3quat.h (exp: Rodrigues + V, Quaterniond(R), operator*, normalizeRotation).  out7 = quaternion x y z w + translation.
void ref_g2o_oplus(const double* pose7, const double* update6, double* out7) {
  g2o::VertexSE3Expmap v;
  v.setEstimate(se3(pose7));
  v.oplus(update6);
  const g2o::SE3Quat& T = v.estimate();
  out7[0] = T.rotation().x(); out7[1] = T.rotation().y(); out7[2] = T.rotation().z(); out7[3] = T.rotation().w();
  out7[4] = T.translation()[0]; out7[5] = T.translation()[1]; out7[6] = T.translation()[2];
}

// SE3Quat::map and SE3Quat::operator* / inverse of the reference's se3quat.h (what the edges call), for direct checks
void ref_g2o_se3_map(const double* pose7, const double* X, double* out3) {
  const Eigen::Vector3d r = se3(pose7).map(Eigen::Vector3d(X[0], X[1], X[2]));
  out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2];
}

}  // extern "C"
