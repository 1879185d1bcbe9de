// Replacement for Optimizer::LocalBundleAdjustment (reference src/Optimizer.cc:
// 1116-1498, signature include/Optimizer.h:57).  Steps 1-4 (local KFs, local
// MapPoints, fixed KFs, edge set-up) and 6-7 (outlier erase, write-back under
// Map::mMutexMapUpdate) are the reference's logic on its own data structures;
// step 5, g2o's optimizer.optimize(10), is lba_solve() of liborbb200.so.  Build
// inside the ORB_SLAM3 tree in place of that one function.  Syntax-checked against the reference's headers
// over stand-ins for its third-party libraries (tests/test_shim_syntax.py); not linked here -- see INTEGRATION.md.
#include <map>
#include <memory>
#include <stdexcept>

#include "Optimizer.h"
#include "orb_b200.h"
#include "orbb200_gate.h"

namespace ORB_SLAM3 {

void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF, int& num_OptKF,
                                      int& num_MPs, int& num_edges) {
  // lba_solve knows both camera models of the reference (Pinhole, KannalaBrandt8) on the mono edges and the
  // second-camera edges (EdgeSE3ProjectXYZToBody, Optimizer.cc:1366-1400); a camera model it does not know hands the
  // whole call to the reference body before any mnBALocalForKF / mnBAFixedForKF mark is written (the reference body
  // writes the same marks itself)
  if (!orbb200_gate::lba_gpu_path(pKF))
    return LocalBundleAdjustment_Reference(pKF, pbStopFlag, pMap, num_fixedKF, num_OptKF, num_MPs, num_edges);
  // ---- 1-3: identical to Optimizer.cc:1119-1186
  list<KeyFrame*> lLocalKeyFrames;
  lLocalKeyFrames.push_back(pKF);
  pKF->mnBALocalForKF = pKF->mnId;
  Map* pCurrentMap = pKF->GetMap();
  for (KeyFrame* pKFi : pKF->GetVectorCovisibleKeyFrames()) {
    pKFi->mnBALocalForKF = pKF->mnId;
    if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) lLocalKeyFrames.push_back(pKFi);
  }
  num_fixedKF = 0;
  list<MapPoint*> lLocalMapPoints;
  for (KeyFrame* pKFi : lLocalKeyFrames) {
    if (pKFi->mnId == pMap->GetInitKFid()) num_fixedKF = 1;
    for (MapPoint* pMP : pKFi->GetMapPointMatches())
      if (pMP && !pMP->isBad() && pMP->GetMap() == pCurrentMap && pMP->mnBALocalForKF != pKF->mnId) {
        lLocalMapPoints.push_back(pMP);
        pMP->mnBALocalForKF = pKF->mnId;
      }
  }
  list<KeyFrame*> lFixedCameras;
  for (MapPoint* pMP : lLocalMapPoints)
    for (const auto& obs : pMP->GetObservations()) {
      KeyFrame* pKFi = obs.first;
      if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
        pKFi->mnBAFixedForKF = pKF->mnId;
        if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) lFixedCameras.push_back(pKFi);
      }
    }
  num_fixedKF = lFixedCameras.size() + num_fixedKF;
  if (num_fixedKF == 0) {
    Verbose::PrintMess("LM-LBA: There are 0 fixed KF in the optimizations, LBA aborted", Verbose::VERBOSITY_NORMAL);
    return;
  }
  // ---- 4: flatten vertices and edges (:1213-1400) into the graph view
  std::map<KeyFrame*, int> kf_index;
  std::vector<KeyFrame*> kfs;
  std::vector<double> kf_pose;
  std::vector<uint8_t> kf_fixed;
  std::vector<float> kf_cam, kf_cam_dist, kf_cam2;
  std::vector<uint8_t> kf_cam_model, kf_cam2_model;
  std::vector<double> kf_trl;
  auto add_kf = [&](KeyFrame* k, bool fixed) {
    kf_index[k] = (int)kfs.size();
    kfs.push_back(k);
    const Sophus::SE3f Tcw = k->GetPose();
    const Eigen::Quaterniond q = Tcw.unit_quaternion().cast<double>();  // g2o::SE3Quat(q, t) (:1217)
    const Eigen::Vector3d t = Tcw.translation().cast<double>();
    const double p[7] = {q.x(), q.y(), q.z(), q.w(), t(0), t(1), t(2)};
    kf_pose.insert(kf_pose.end(), p, p + 7);
    kf_fixed.push_back(fixed ? 1 : 0);
    // rig fields: e->pCamera = pKFi->mpCamera of the mono edges (:1326), mpCamera2 + Trl of the body edges (:1384-1387)
    const bool kb8 = k->mpCamera->GetType() == GeometricCamera::CAM_FISHEYE;
    kf_cam_model.push_back(kb8 ? ORB_CAM_KB8 : ORB_CAM_PINHOLE);
    // stereo edges read the keyframe's fx.. (:1352-1356), mono edges project through mpCamera: one rig, same numbers
    const float c[5] = {kb8 ? k->mpCamera->getParameter(0) : k->fx, kb8 ? k->mpCamera->getParameter(1) : k->fy,
                        kb8 ? k->mpCamera->getParameter(2) : k->cx, kb8 ? k->mpCamera->getParameter(3) : k->cy, k->mbf};
    kf_cam.insert(kf_cam.end(), c, c + 5);
    for (int i = 4; i < 8; i++) kf_cam_dist.push_back(kb8 ? k->mpCamera->getParameter(i) : 0.f);
    if (k->mpCamera2) {
      const bool kb8r = k->mpCamera2->GetType() == GeometricCamera::CAM_FISHEYE;
      kf_cam2_model.push_back(kb8r ? ORB_CAM_KB8 : ORB_CAM_PINHOLE);
      for (int i = 0; i < 8; i++) kf_cam2.push_back(i < 4 || kb8r ? k->mpCamera2->getParameter(i) : 0.f);
      const Sophus::SE3f Trl = k->GetRelativePoseTrl();
      const Eigen::Quaterniond qr = Trl.unit_quaternion().cast<double>();
      const Eigen::Vector3d tr = Trl.translation().cast<double>();
      const double r[7] = {qr.x(), qr.y(), qr.z(), qr.w(), tr(0), tr(1), tr(2)};
      kf_trl.insert(kf_trl.end(), r, r + 7);
    } else {
      kf_cam2_model.push_back(ORB_CAM_PINHOLE);
      kf_cam2.insert(kf_cam2.end(), 8, 0.f);
      const double r[7] = {0, 0, 0, 1, 0, 0, 0};
      kf_trl.insert(kf_trl.end(), r, r + 7);
    }
  };
  for (KeyFrame* k : lLocalKeyFrames) add_kf(k, k->mnId == pMap->GetInitKFid());
  for (KeyFrame* k : lFixedCameras) add_kf(k, true);
  num_OptKF = lLocalKeyFrames.size();
  std::vector<MapPoint*> mps(lLocalMapPoints.begin(), lLocalMapPoints.end());
  std::vector<double> mp_pos;
  std::vector<int32_t> e_kf, e_mp;
  std::vector<uint8_t> e_stereo;
  std::vector<double> e_obs;
  std::vector<float> e_is2;
  std::vector<KeyFrame*> e_kfptr;
  for (size_t l = 0; l < mps.size(); l++) {
    const Eigen::Vector3d X = mps[l]->GetWorldPos().cast<double>();
    mp_pos.insert(mp_pos.end(), {X(0), X(1), X(2)});
    for (const auto& obs : mps[l]->GetObservations()) {
      KeyFrame* pKFi = obs.first;
      if (pKFi->isBad() || pKFi->GetMap() != pCurrentMap) continue;
      const int leftIndex = get<0>(obs.second);
      if (leftIndex != -1) {
        const cv::KeyPoint& kpUn = pKFi->mvKeysUn[leftIndex];
        const float ur = pKFi->mvuRight[leftIndex];
        e_kf.push_back(kf_index.at(pKFi));
        e_mp.push_back((int32_t)l);
        e_stereo.push_back(ur >= 0 ? LBA_EDGE_STEREO : LBA_EDGE_MONO);   // :1305 / :1332
        e_obs.insert(e_obs.end(), {(double)kpUn.pt.x, (double)kpUn.pt.y, (double)ur});
        e_is2.push_back(pKFi->mvInvLevelSigma2[kpUn.octave]);     // :1316 / :1345
        e_kfptr.push_back(pKFi);
      }
      if (pKFi->mpCamera2) {  // :1366-1400: the observation of the second camera, an edge to the same (body) pose
        int rightIndex = get<1>(obs.second);
        if (rightIndex != -1) {
          rightIndex -= pKFi->NLeft;
          const cv::KeyPoint& kp = pKFi->mvKeysRight[rightIndex];
          e_kf.push_back(kf_index.at(pKFi));
          e_mp.push_back((int32_t)l);
          e_stereo.push_back(LBA_EDGE_BODY);
          e_obs.insert(e_obs.end(), {(double)kp.pt.x, (double)kp.pt.y, -1.0});
          e_is2.push_back(pKFi->mvInvLevelSigma2[kp.octave]);     // :1378
          e_kfptr.push_back(pKFi);
        }
      }
    }
  }
  num_MPs = mps.size();
  num_edges = e_kf.size();
  if (pbStopFlag && *pbStopFlag) return;  // :1406-1408
  // ---- 5: optimizer.optimize(10) on the GPU
  lba_graph_view g{(int32_t)kfs.size(), kf_pose.data(), kf_fixed.data(), kf_cam.data(), (int32_t)mps.size(),
                   mp_pos.data(), (int32_t)e_kf.size(), e_kf.data(), e_mp.data(), e_stereo.data(), e_obs.data(),
                   e_is2.data(), kf_cam_model.data(), kf_cam_dist.data(), kf_cam2_model.data(), kf_cam2.data(),
                   kf_trl.data()};
  thread_local std::unique_ptr<lba_solver, void (*)(lba_solver*)> solver(nullptr, lba_destroy);
  if (!solver) {
    lba_solver* s = nullptr;
    if (lba_create(0, &s) != ORB_OK) throw std::runtime_error(orb_last_error());
    solver.reset(s);
  }
  std::vector<double> pose_out(kf_pose.size()), pos_out(mp_pos.size()), chi2(e_kf.size());
  std::vector<uint8_t> depth_pos(e_kf.size());
  lba_stats stats;
  const double lambda_init = pMap->IsInertial() ? 100.0 : 0.0;  // :1197-1198
  if (lba_solve(solver.get(), &g, reinterpret_cast<const volatile uint8_t*>(pbStopFlag), 10, lambda_init,
                pose_out.data(), pos_out.data(), chi2.data(), depth_pos.data(), &stats) < 0)
    throw std::runtime_error(orb_last_error());
  // ---- 6: chi2 / depth test (:1413-1460)
  vector<pair<KeyFrame*, MapPoint*> > vToErase;
  for (size_t e = 0; e < e_kf.size(); e++) {
    MapPoint* pMP = mps[e_mp[e]];
    if (pMP->isBad()) continue;
    // mono :1424, second camera :1438 (5.991), stereo :1453 (7.815)
    if (chi2[e] > (e_stereo[e] == LBA_EDGE_STEREO ? 7.815 : 5.991) || !depth_pos[e]) vToErase.push_back(make_pair(e_kfptr[e], pMP));
  }
  // ---- 7: write-back under the map mutex (:1464-1497)
  unique_lock<mutex> lock(pMap->mMutexMapUpdate);
  for (auto& er : vToErase) {
    er.first->EraseMapPointMatch(er.second);
    er.second->EraseObservation(er.first);
  }
  for (KeyFrame* k : lLocalKeyFrames) {
    const double* p = &pose_out[7 * kf_index.at(k)];
    k->SetPose(Sophus::SE3f(Eigen::Quaterniond(p[3], p[0], p[1], p[2]).cast<float>(),
                            Eigen::Vector3d(p[4], p[5], p[6]).cast<float>()));
  }
  for (size_t l = 0; l < mps.size(); l++) {
    mps[l]->SetWorldPos(Eigen::Vector3d(pos_out[3 * l], pos_out[3 * l + 1], pos_out[3 * l + 2]).cast<float>());
    mps[l]->UpdateNormalAndDepth();
  }
  pMap->IncreaseChangeIndex();
}

}  // namespace ORB_SLAM3
