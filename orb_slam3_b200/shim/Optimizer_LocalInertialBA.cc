// Replacement for Optimizer::LocalInertialBA (reference src/Optimizer.cc:2383-2958).
// Window selection (:2385-2497), vertex / edge bookkeeping (:2523-2735), outlier erasure and write-back
// (:2755-2868) are the reference's logic on its own data structures, restated over index lists; the
// g2o::SparseOptimizer in the middle -- initializeOptimization / computeActiveErrors / optimize(opt_it)
// (:2748-2751) -- is lia_solve() of liborbb200.so on the flat graph below.  Rigs or modes the engine does not
// carry (a second camera, a KannalaBrandt8 camera, bRecInit's robust kernel on every inertial edge) call
// LocalInertialBA_Reference, the reference body kept under that name.
// Syntax-checked against the reference's headers over stand-ins for its third-party libraries
// (tests/test_shim_syntax.py); the device path behind lia_solve is validated on the B200 against the oracle
// (tests/test_lia_gpu.py).
#include <cmath>
#include <iostream>
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "Optimizer.h"
#include "orb_b200.h"
#include "orbb200_gate.h"

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

void Optimizer::LocalInertialBA(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF, int& num_OptKF, int& num_MPs,
                                int& num_edges, bool bLarge, bool bRecInit) {
  // lia_solve carries one Pinhole camera per keyframe and the robust kernel of the window's last inertial edge only
  if (!orbb200_gate::gpu_path(pKF) || bRecInit)
    return LocalInertialBA_Reference(pKF, pbStopFlag, pMap, num_fixedKF, num_OptKF, num_MPs, num_edges, bLarge, bRecInit);
  Map* pCurrentMap = pKF->GetMap();
  const int maxOpt = bLarge ? 25 : 10;                                      // :2387-2392
  const int Nd = std::min((int)pCurrentMap->KeyFramesInMap() - 2, maxOpt);
  // ---- temporal window: pKF and its predecessors (:2395-2412)
  std::vector<KeyFrame*> opt;
  opt.reserve(Nd);
  opt.push_back(pKF);
  pKF->mnBALocalForKF = pKF->mnId;
  for (int i = 1; i < Nd && opt.back()->mPrevKF; i++) {
    opt.push_back(opt.back()->mPrevKF);
    opt.back()->mnBALocalForKF = pKF->mnId;
  }
  // ---- map points of the window (:2416-2433)
  std::vector<MapPoint*> mps;
  for (KeyFrame* k : opt)
    for (MapPoint* pMP : k->GetMapPointMatches())
      if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) {
        mps.push_back(pMP);
        pMP->mnBALocalForKF = pKF->mnId;
      }
  // ---- fixed keyframes: the one before the window, or the window's oldest (:2436-2448); then, per map point, the
  //      first observer outside the window (:2483-2503).  (The covisible optimisable set of :2451-2480 is empty in
  //      the reference: maxCovKF = 0.)
  std::list<KeyFrame*> fixedKFs;
  if (opt.back()->mPrevKF) {
    fixedKFs.push_back(opt.back()->mPrevKF);
    opt.back()->mPrevKF->mnBAFixedForKF = pKF->mnId;
  } else {
    opt.back()->mnBALocalForKF = 0;
    opt.back()->mnBAFixedForKF = pKF->mnId;
    fixedKFs.push_back(opt.back());
    opt.pop_back();
  }
  const size_t maxFixKF = 200;
  for (MapPoint* pMP : mps) {
    for (const auto& ob : pMP->GetObservations()) {
      KeyFrame* pKFi = ob.first;
      if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
        pKFi->mnBAFixedForKF = pKF->mnId;
        if (!pKFi->isBad()) { fixedKFs.push_back(pKFi); break; }
      }
    }
    if (fixedKFs.size() >= maxFixKF) break;
  }
  // ---- vertices (:2523-2561): window first (newest first), then the fixed keyframes
  const int N = (int)opt.size();
  LiaFlatGraph G;
  std::map<KeyFrame*, int> kf_index;
  for (KeyFrame* k : opt) { kf_index[k] = (int)kf_index.size(); LiaAddKeyFrame(G, k, false); }
  for (KeyFrame* k : fixedKFs) { kf_index[k] = (int)kf_index.size(); LiaAddKeyFrame(G, k, true); }
  // ---- inertial edges (:2564-2622): keyframe i of the window and its predecessor
  for (int i = 0; i < N; i++) {
    KeyFrame* k = opt[i];
    if (!k->mPrevKF) { std::cout << "NOT INERTIAL LINK TO PREVIOUS FRAME!!!!" << std::endl; continue; }
    if (k->bImu && k->mPrevKF->bImu && k->mpImuPreintegrated) {
      k->mpImuPreintegrated->SetNewBias(k->mPrevKF->GetImuBias());
      auto prev = kf_index.find(k->mPrevKF);
      if (prev == kf_index.end()) { std::cerr << "Error: previous keyframe is not a vertex" << std::endl; continue; }
      LiaAddInertial(G, prev->second, i, k->mpImuPreintegrated, i == N - 1);
    } else {
      std::cout << "ERROR building inertial edge" << std::endl;
    }
  }
  // ---- map points and visual edges (:2660-2735); a single-camera rig has no right-index observations
  std::vector<KeyFrame*> e_kfptr;
  for (size_t l = 0; l < mps.size(); l++) {
    push3(G.mp, mps[l]->GetWorldPos().cast<double>());
    for (const auto& ob : mps[l]->GetObservations()) {
      KeyFrame* pKFi = ob.first;
      if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) continue;
      if (pKFi->isBad() || pKFi->GetMap() != pCurrentMap) continue;
      const int leftIndex = std::get<0>(ob.second);
      if (leftIndex == -1) continue;
      const cv::KeyPoint& kpUn = pKFi->mvKeysUn[leftIndex];
      const float ur = pKFi->mvuRight[leftIndex];
      Eigen::Matrix<double, 2, 1> uv;
      uv << kpUn.pt.x, kpUn.pt.y;
      const float unc2 = pKFi->mpCamera->uncertainty2(uv);            // :2688 / :2714
      G.e_kf.push_back(kf_index.at(pKFi));
      G.e_mp.push_back((int32_t)l);
      G.stereo.push_back(ur >= 0 ? 1 : 0);
      G.obs.insert(G.obs.end(), {(double)kpUn.pt.x, (double)kpUn.pt.y, (double)ur});
      G.is2.push_back(pKFi->mvInvLevelSigma2[kpUn.octave] / unc2);
      e_kfptr.push_back(pKFi);
    }
  }
  num_fixedKF = (int)fixedKFs.size(); num_OptKF = N; num_MPs = (int)mps.size(); num_edges = (int)G.e_kf.size();
  // ---- optimize(opt_it) on the GPU
  std::vector<double> kf_out, mp_out, chi2;
  std::vector<uint8_t> depth_pos;
  double stats[8];
  LiaSolve(G, pKF, bLarge, kf_out, mp_out, chi2, depth_pos, stats);
  const float err = (float)stats[2], err_end = (float)stats[3];          // float in the reference (:2750-2752)
  // ---- inlier test (:2760-2790)
  std::vector<std::pair<KeyFrame*, MapPoint*> > vToErase;
  for (size_t e = 0; e < G.e_kf.size(); e++) {
    MapPoint* pMP = mps[G.e_mp[e]];
    if (pMP->isBad()) continue;
    bool bad;
    if (G.stereo[e]) {
      bad = chi2[e] > 7.815;
    } else {
      const bool bClose = pMP->mTrackDepth < 10.f;
      bad = (chi2[e] > 5.991 && !bClose) || (chi2[e] > 1.5f * 5.991f && bClose) || !depth_pos[e];
    }
    if (bad) vToErase.push_back(std::make_pair(e_kfptr[e], pMP));
  }
  // ---- write-back under the map mutex (:2793-2868)
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
  if ((2 * err < err_end || std::isnan(err) || std::isnan(err_end)) && !bLarge) {
    std::cout << "FAIL LOCAL-INERTIAL BA!!!!" << std::endl;
    return;
  }
  for (auto& er : vToErase) {
    er.first->EraseMapPointMatch(er.second);
    er.second->EraseObservation(er.first);
  }
  for (KeyFrame* k : fixedKFs) k->mnBAFixedForKF = 0;
  for (int i = 0; i < N; i++) {
    KeyFrame* k = opt[i];
    const double* o = &kf_out[21 * (size_t)i];
    Eigen::Matrix3d Rcw;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rcw(r, c) = o[3 * r + c];
    k->SetPose(Sophus::SE3f(Rcw.cast<float>(), Eigen::Vector3d(o[9], o[10], o[11]).cast<float>()));
    k->mnBALocalForKF = 0;
    if (k->bImu) {
      k->SetVelocity(Eigen::Vector3d(o[12], o[13], o[14]).cast<float>());
      // IMU::Bias(b_acc_x, b_acc_y, b_acc_z, b_ang_vel_x, b_ang_vel_y, b_ang_vel_z) from (gyro, acc) (:2846-2848)
      k->SetNewBias(IMU::Bias((float)o[18], (float)o[19], (float)o[20], (float)o[15], (float)o[16], (float)o[17]));
    }
  }
  for (size_t l = 0; l < mps.size(); l++) {
    mps[l]->SetWorldPos(Eigen::Vector3d(mp_out[3 * l], mp_out[3 * l + 1], mp_out[3 * l + 2]).cast<float>());
    mps[l]->UpdateNormalAndDepth();
  }
  pMap->IncreaseChangeIndex();
}

}  // namespace ORB_SLAM3
