// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's visual-inertial graph types as object code
// (oracle/_ref/libref_vi.so): /root/reference/src/G2oTypes.cc and src/CameraModels/Pinhole.cpp compiled UNMODIFIED against
// oracle/eigencompat/ (functional fixed-size Eigen incl. blocks, g2o vertex / edge bases) and oracle/cvcompat/; everything
// else they name is bound to stubs that print the symbol and abort (oracle/Makefile).  An ImuCamPose is filled member by
// member (they are public) from a lia_graph_view and the reference's own
//   EdgeMono / EdgeStereo ::computeError, ::linearizeOplus, ::isDepthPositive   (include/G2oTypes.h:342-452, src/G2oTypes.cc:290-345)
//   ImuCamPose::Project / ProjectStereo / isDepthPositive                        (src/G2oTypes.cc:167-190)
// run as object code: the visual edges of Optimizer::LocalInertialBA (Optimizer.cc:2636-2735).  The inertial edges of the
// same file (EdgeInertial, EdgeGyroRW, EdgeAccRW) need dynamic-size Jacobian blocks, an eigen-decomposition and an SVD of
// the real Eigen and are NOT compared.  tests/test_ref_edges.py holds the oracle (orc_lia.cpp) against these.
// Nothing in the product links this.
#include <vector>

#include "G2oTypes.h"
#include "Pinhole.h"
#include "../include/orb_b200.h"

using namespace ORB_SLAM3;

namespace {
Eigen::Matrix3d mat3(const double* p) { Eigen::Matrix3d M; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M(r, c) = p[3 * r + c]; return M; }
Eigen::Vector3d vec3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
template <class M> void dump(const M& J, int rows, int cols, double* out) { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) out[r * cols + c] = J(r, c); }
}  // namespace

extern "C" int ref_vi_edge(const lia_graph_view* g, int e, double* err3, double* A9, double* B18, int* depth_pos) {
  if (!g || e < 0 || e >= g->n_edges) return -1;
  const int k = g->e_kf[e], l = g->e_mp[e];
  Pinhole cam(std::vector<float>{g->fx, g->fy, g->cx, g->cy});
  ImuCamPose P;   // what ImuCamPose(KeyFrame*) sets (G2oTypes.cc:23-61), one camera
  P.its = 0;
  P.Rwb = mat3(g->kf_Rwb + 9 * (size_t)k); P.twb = vec3(g->kf_twb + 3 * (size_t)k);
  P.Rcw.push_back(mat3(g->kf_Rcw + 9 * (size_t)k)); P.tcw.push_back(vec3(g->kf_tcw + 3 * (size_t)k));
  P.Rcb.push_back(mat3(g->Rcb)); P.tcb.push_back(vec3(g->tcb));
  P.Rbc.push_back(P.Rcb[0].transpose()); P.tbc.push_back(vec3(g->tbc));
  P.pCamera.push_back(&cam);
  P.bf = g->bf;
  P.Rwb0 = P.Rwb; P.DR.setIdentity();
  VertexPose vp;
  vp.setEstimate(P);
  g2o::VertexSBAPointXYZ vx;
  vx.setEstimate(vec3(g->mp_pos + 3 * (size_t)l));
  const double* o = g->e_obs + 3 * (size_t)e;
  for (int i = 0; i < 9; i++) A9[i] = 0;
  for (int i = 0; i < 18; i++) B18[i] = 0;
  err3[2] = 0;
  if (g->e_stereo[e]) {
    EdgeStereo ed(0);
    ed.setVertex(0, &vx); ed.setVertex(1, &vp);
    ed.setMeasurement(Eigen::Vector3d(o[0], o[1], o[2]));
    ed.computeError(); ed.linearizeOplus();
    for (int i = 0; i < 3; i++) err3[i] = ed.error()[i];
    dump(ed.jacobianOplusXi(), 3, 3, A9); dump(ed.jacobianOplusXj(), 3, 6, B18);
    // (EdgeStereo has no isDepthPositive; the optimiser never asks: Optimizer.cc:2809-2836 tests chi2 only)
    *depth_pos = vp.estimate().isDepthPositive(vx.estimate(), 0);
  } else {
    EdgeMono ed(0);
    ed.setVertex(0, &vx); ed.setVertex(1, &vp);
    ed.setMeasurement(Eigen::Vector2d(o[0], o[1]));
    ed.computeError(); ed.linearizeOplus();
    for (int i = 0; i < 2; i++) err3[i] = ed.error()[i];
    dump(ed.jacobianOplusXi(), 2, 3, A9); dump(ed.jacobianOplusXj(), 2, 6, B18);
    *depth_pos = ed.isDepthPositive();
  }
  return 0;
}
