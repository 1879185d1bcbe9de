// Replacement for Optimizer::LocalInertialBA (reference src/Optimizer.cc:2383-2958).
// The window selection, vertex / edge bookkeeping, outlier erasure and write-back are the reference's
// logic on its own data structures (kept verbatim in the tree under `#ifndef ORB_B200_HOTPATH` ordering,
// see INTEGRATION.md); this unit shows only the part that changes: the flat graph handed to lia_solve()
// instead of building a g2o::SparseOptimizer, and how the outputs map back.
// Syntax-checked against the reference's headers over stand-ins for its third-party libraries (tests/test_shim_syntax.py).  The device path behind
// lia_solve is validated on the B200 against the oracle (tests/test_lia_gpu.py); this unit is the graph
// flattening only, not a finished replacement of the function.
#include <stdexcept>
#include <string>
#include <vector>

#include "Optimizer.h"
#include "orb_b200.h"

namespace ORB_SLAM3 {

// vKFs: vpOptimizableKFs (newest first) followed by lFixedKeyFrames; vMPs: lLocalMapPoints;
// the edge lists are what the reference loops :2636-2735 enumerate, with indices into vKFs / vMPs.
struct LiaFlatGraph {
  std::vector<double> Rwb, twb, Rcw, tcw, vel, bg, ba, mp, obs;
  std::vector<uint8_t> fixed, has_imu, stereo, last;
  std::vector<int32_t> e_kf, e_mp, i_kf1, i_kf2;
  std::vector<float> is2, dR, dV, dP, JRg, JVg, JVa, JPg, JPa, bias, dT, C;
};

static void push3(std::vector<double>& v, const Eigen::Vector3d& x) { v.insert(v.end(), {x[0], x[1], x[2]}); }
static void push9(std::vector<double>& v, const Eigen::Matrix3d& M) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) v.push_back(M(r, c));
}
static void push9f(std::vector<float>& v, const Eigen::Matrix3f& M) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) v.push_back(M(r, c));
}

// One keyframe = one VertexPose (+ VertexVelocity / VertexGyroBias / VertexAccBias when bImu): the values the
// ImuCamPose / vertex constructors read (G2oTypes.cc:25-70, G2oTypes.h:191-260).
void LiaAddKeyFrame(LiaFlatGraph& G, KeyFrame* pKF, bool fixed) {
  push9(G.Rwb, pKF->GetImuRotation().cast<double>());
  push3(G.twb, pKF->GetImuPosition().cast<double>());
  push9(G.Rcw, pKF->GetRotation().cast<double>());
  push3(G.tcw, pKF->GetTranslation().cast<double>());
  push3(G.vel, pKF->GetVelocity().cast<double>());
  push3(G.bg, pKF->GetGyroBias().cast<double>());
  push3(G.ba, pKF->GetAccBias().cast<double>());
  G.fixed.push_back(fixed);
  G.has_imu.push_back(pKF->bImu);
}

// One EdgeInertial + EdgeGyroRW + EdgeAccRW (:2580-2611); pInt->SetNewBias(prev bias) has been called (:2569)
void LiaAddInertial(LiaFlatGraph& G, int kf1, int kf2, IMU::Preintegrated* pInt, bool last) {
  G.i_kf1.push_back(kf1); G.i_kf2.push_back(kf2); G.last.push_back(last);
  push9f(G.dR, pInt->dR);
  for (int c = 0; c < 3; c++) { G.dV.push_back(pInt->dV[c]); G.dP.push_back(pInt->dP[c]); }
  push9f(G.JRg, pInt->JRg); push9f(G.JVg, pInt->JVg); push9f(G.JVa, pInt->JVa); push9f(G.JPg, pInt->JPg); push9f(G.JPa, pInt->JPa);
  const IMU::Bias b = pInt->GetOriginalBias();
  G.bias.insert(G.bias.end(), {b.bax, b.bay, b.baz, b.bwx, b.bwy, b.bwz});
  G.dT.push_back(pInt->dT);
  for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) G.C.push_back(pInt->C(r, c));
}

// optimizer.initializeOptimization(); computeActiveErrors(); optimize(opt_it) (:2748-2751) -> the engine
void LiaSolve(const LiaFlatGraph& G, KeyFrame* pKF, bool bLarge, std::vector<double>& kf_out, std::vector<double>& mp_out,
              std::vector<double>& chi2, std::vector<uint8_t>& depth_pos, double stats[8]) {
  lia_graph_view v;
  v.n_kf = (int32_t)G.fixed.size();
  v.kf_Rwb = G.Rwb.data(); v.kf_twb = G.twb.data(); v.kf_Rcw = G.Rcw.data(); v.kf_tcw = G.tcw.data();
  v.kf_fixed = G.fixed.data(); v.kf_has_imu = G.has_imu.data(); v.kf_vel = G.vel.data(); v.kf_bg = G.bg.data(); v.kf_ba = G.ba.data();
  const Eigen::Matrix3d Rcb = pKF->mImuCalib.mTcb.rotationMatrix().cast<double>();
  const Eigen::Vector3d tcb = pKF->mImuCalib.mTcb.translation().cast<double>(), tbc = pKF->mImuCalib.mTbc.translation().cast<double>();
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) v.Rcb[3 * r + c] = Rcb(r, c); v.tcb[r] = tcb[r]; v.tbc[r] = tbc[r]; }
  v.fx = pKF->fx; v.fy = pKF->fy; v.cx = pKF->cx; v.cy = pKF->cy; v.bf = pKF->mbf;
  v.n_mp = (int32_t)(G.mp.size() / 3); v.mp_pos = G.mp.data();
  v.n_edges = (int32_t)G.e_kf.size(); v.e_kf = G.e_kf.data(); v.e_mp = G.e_mp.data(); v.e_stereo = G.stereo.data();
  v.e_obs = G.obs.data(); v.e_inv_sigma2 = G.is2.data();
  v.n_inertial = (int32_t)G.i_kf1.size(); v.i_kf1 = G.i_kf1.data(); v.i_kf2 = G.i_kf2.data();
  v.i_dR = G.dR.data(); v.i_dV = G.dV.data(); v.i_dP = G.dP.data(); v.i_JRg = G.JRg.data(); v.i_JVg = G.JVg.data();
  v.i_JVa = G.JVa.data(); v.i_JPg = G.JPg.data(); v.i_JPa = G.JPa.data(); v.i_bias = G.bias.data(); v.i_dT = G.dT.data();
  v.i_C = G.C.data(); v.i_last = G.last.data();
  v.lambda_init = bLarge ? 1e-2 : 1e0;   // :2509-2519
  v.iterations = bLarge ? 4 : 10;        // :2387-2392
  static thread_local orb_lia* h = nullptr;  // LocalMapping thread
  if (!h && lia_create(/*device=*/0, &h) != ORB_OK) throw std::runtime_error(std::string("lia_create: ") + orb_last_error());
  kf_out.resize(21 * (size_t)v.n_kf); mp_out.resize(3 * (size_t)v.n_mp); chi2.resize(v.n_edges); depth_pos.resize(v.n_edges);
  if (lia_solve(h, &v, kf_out.data(), mp_out.data(), chi2.data(), depth_pos.data(), stats) < 0)
    throw std::runtime_error(std::string("lia_solve: ") + orb_last_error());
  // stats[2] / stats[3] are `err` / `err_end` of the "FAIL LOCAL-INERTIAL BA" test (:2795); chi2 / depth_pos feed the
  // outlier tests (:2760-2790); kf_out rows give Rcw | tcw (SetPose), velocity, gyro bias, acc bias (SetNewBias).
}

}  // namespace ORB_SLAM3
