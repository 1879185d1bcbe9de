// TEST INFRASTRUCTURE ONLY -- appended (oracle/Makefile) to the reference's own robust_kernel.h / robust_kernel_impl.h,
// types_six_dof_expmap.h and the constructQuadraticForm() definitions of core/base_binary_edge.hpp and base_unary_edge.hpp
// (those two sliced from where they lie: the template header line of ::constructQuadraticForm() up to the next member),
// all piped UNMODIFIED.  Builds the g2o graph of a Pinhole LocalBundleAdjustment window the way Optimizer.cc:1142-1362 does
// -- VertexSE3Expmap per keyframe, VertexSBAPointXYZ per landmark, EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ with
// information = invSigma2 * I and a RobustKernelHuber(sqrt(5.991) / sqrt(7.815)) -- and runs, edge by edge in edge order,
// computeError(), linearizeOplus() and constructQuadraticForm(): the accumulation of every edge into the Hessian blocks
// and the right-hand side (row a18) as the reference's object code.  tests/test_ref_edges.py holds the oracle's
// build_system() against the result.  Nothing in the product links this.
#include <cmath>
#include <vector>

#include "../../../../include/orb_b200.h"

namespace {
g2o::SE3Quat se3q(const double* p) {
  return g2o::SE3Quat(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
template <class E> void accumulate(E& e, int d, double* W18) {
  e.computeError();
  e.linearizeOplus();
  e.constructQuadraticForm();
  // _hessian = A^T wOmega B (3 x 6, landmark rows); W of the oracle = its transpose (6 x 3 row-major)
  for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) W18[i * 3 + j] = e.hessianBlock()(j, i);
  (void)d;
}
}  // namespace

// Hpp: n_kf x 36 (row-major 6x6 per keyframe, zero for fixed ones), Hll: n_mp x 9, W: n_edges x 18, bp: n_kf x 6, bl: n_mp x 3
extern "C" int ref_g2o_build_system(const lba_graph_view* g, double* Hpp, double* Hll, double* W, double* bp, double* bl) {
  std::vector<g2o::VertexSE3Expmap> vkf(g->n_kf);
  std::vector<g2o::VertexSBAPointXYZ> vmp(g->n_mp);
  for (int k = 0; k < g->n_kf; k++) { vkf[k].setEstimate(se3q(g->kf_pose + 7 * k)); vkf[k].setFixed(g->kf_fixed[k] != 0); vkf[k].clearQuadraticForm(); }
  for (int l = 0; l < g->n_mp; l++) { vmp[l].setEstimate(Eigen::Vector3d(g->mp_pos[3 * l], g->mp_pos[3 * l + 1], g->mp_pos[3 * l + 2])); vmp[l].clearQuadraticForm(); }
  const float thHuberMono = sqrt(5.991), thHuberStereo = sqrt(7.815);   // Optimizer.cc:1275-1276
  g2o::RobustKernelHuber rk_mono, rk_stereo;
  rk_mono.setDelta(thHuberMono); rk_stereo.setDelta(thHuberStereo);
  for (int e = 0; e < g->n_edges; e++) {
    const int k = g->e_kf[e], l = g->e_mp[e];
    const float* cam = g->kf_cam + 5 * (size_t)k;
    const double* obs = g->e_obs + 3 * (size_t)e;
    const double invSigma2 = g->e_inv_sigma2[e];
    if (g->e_stereo[e] == LBA_EDGE_STEREO) {
      g2o::EdgeStereoSE3ProjectXYZ ed;
      ed.setVertex(0, &vmp[l]); ed.setVertex(1, &vkf[k]);
      ed.setMeasurement(Eigen::Vector3d(obs[0], obs[1], obs[2]));
      ed.setInformation(Eigen::Matrix3d::Identity() * invSigma2);
      ed.setRobustKernel(&rk_stereo);
      ed.fx = cam[0]; ed.fy = cam[1]; ed.cx = cam[2]; ed.cy = cam[3]; ed.bf = cam[4];
      accumulate(ed, 3, W + 18 * (size_t)e);
    } else if (g->e_stereo[e] == LBA_EDGE_MONO) {
      g2o::EdgeSE3ProjectXYZ ed;
      ed.setVertex(0, &vmp[l]); ed.setVertex(1, &vkf[k]);
      ed.setMeasurement(Eigen::Vector2d(obs[0], obs[1]));
      ed.setInformation(Eigen::Matrix2d::Identity() * invSigma2);
      ed.setRobustKernel(&rk_mono);
      ed.fx = cam[0]; ed.fy = cam[1]; ed.cx = cam[2]; ed.cy = cam[3];
      accumulate(ed, 2, W + 18 * (size_t)e);
    } else {
      return -1;   // second-camera edges are ORB_SLAM3's own type (OptimizableTypes.cpp), not g2o's
    }
  }
  for (int k = 0; k < g->n_kf; k++)
    for (int i = 0; i < 6; i++) { bp[6 * k + i] = vkf[k].b()[i]; for (int j = 0; j < 6; j++) Hpp[36 * (size_t)k + i * 6 + j] = vkf[k].A()(i, j); }
  for (int l = 0; l < g->n_mp; l++)
    for (int i = 0; i < 3; i++) { bl[3 * l + i] = vmp[l].b()[i]; for (int j = 0; j < 3; j++) Hll[9 * (size_t)l + i * 3 + j] = vmp[l].A()(i, j); }
  return 0;
}

// The frame graph of Optimizer::PoseOptimization at its first linearisation (Optimizer.cc:851-935): one VertexSE3Expmap at
// the frame pose, EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose per matched point with information =
// invSigma2 * I and RobustKernelHuber(sqrt(5.991) / sqrt(7.815)); computeError, linearizeOplus and BaseUnaryEdge's
// constructQuadraticForm edge by edge.  H36 row-major, b6.
extern "C" int ref_g2o_pose_system(const pose_opt_view* v, double* H36, double* b6) {
  g2o::VertexSE3Expmap vpose;
  const double p7[7] = {v->pose[0], v->pose[1], v->pose[2], v->pose[3], v->pose[4], v->pose[5], v->pose[6]};
  vpose.setEstimate(se3q(p7));
  vpose.setFixed(false);
  vpose.clearQuadraticForm();
  const float deltaMono = sqrt(5.991), deltaStereo = sqrt(7.815);   // :851-852
  g2o::RobustKernelHuber rk_mono, rk_stereo;
  rk_mono.setDelta(deltaMono); rk_stereo.setDelta(deltaStereo);
  for (int e = 0; e < v->n; e++) {
    const Eigen::Vector3d Xw((double)v->xw[3 * e], (double)v->xw[3 * e + 1], (double)v->xw[3 * e + 2]);   // GetWorldPos().cast<double>()
    const double invSigma2 = v->inv_sigma2[e];
    if (v->obs[3 * e + 2] >= 0) {
      g2o::EdgeStereoSE3ProjectXYZOnlyPose ed;
      ed.setVertex(0, &vpose);
      ed.setMeasurement(Eigen::Vector3d((double)v->obs[3 * e], (double)v->obs[3 * e + 1], (double)v->obs[3 * e + 2]));
      ed.setInformation(Eigen::Matrix3d::Identity() * invSigma2);
      ed.setRobustKernel(&rk_stereo);
      ed.fx = v->fx; ed.fy = v->fy; ed.cx = v->cx; ed.cy = v->cy; ed.bf = v->bf; ed.Xw = Xw;
      ed.computeError(); ed.linearizeOplus(); ed.constructQuadraticForm();
    } else {
      g2o::EdgeSE3ProjectXYZOnlyPose ed;
      ed.setVertex(0, &vpose);
      ed.setMeasurement(Eigen::Vector2d((double)v->obs[3 * e], (double)v->obs[3 * e + 1]));
      ed.setInformation(Eigen::Matrix2d::Identity() * invSigma2);
      ed.setRobustKernel(&rk_mono);
      ed.fx = v->fx; ed.fy = v->fy; ed.cx = v->cx; ed.cy = v->cy; ed.Xw = Xw;
      ed.computeError(); ed.linearizeOplus(); ed.constructQuadraticForm();
    }
  }
  for (int i = 0; i < 6; i++) { b6[i] = vpose.b()[i]; for (int j = 0; j < 6; j++) H36[i * 6 + j] = vpose.A()(i, j); }
  return 0;
}
