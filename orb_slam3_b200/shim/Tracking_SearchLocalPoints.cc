// Replacement for Tracking::SearchLocalPoints (reference src/Tracking.cc:3343-3423): the per-point
// calls of Frame::isInFrustum (src/Frame.cc:512-570) become ONE
// frame_is_in_frustum() over all candidate local map points; the members the reference function
// writes on each MapPoint are written back from the flat outputs.  The fisheye-stereo rig
// (Frame::Nleft != -1) keeps the reference loop.  Everything else (visibility counters, thresholds,
// the SearchByProjection call, which shim/ORBmatcher_hotpath.cc already routes to the engine) is the
// reference's own logic.  Syntax-checked against the reference's headers over stand-ins for its third-party libraries (tests/test_shim_syntax.py); not linked here -- see INTEGRATION.md.
#include <stdexcept>
#include <string>
#include <vector>

#include "Tracking.h"
#include "ORBmatcher.h"
#include "orb_b200.h"
#include "orbb200_gate.h"

namespace ORB_SLAM3 {

void Tracking::SearchLocalPoints() {
  if (!orbb200_gate::gpu_path(mCurrentFrame)) { SearchLocalPoints_Reference(); return; }
  // :3345-3363, unchanged
  for (auto vit = mCurrentFrame.mvpMapPoints.begin(), vend = mCurrentFrame.mvpMapPoints.end(); vit != vend; vit++) {
    MapPoint* pMP = *vit;
    if (!pMP) continue;
    if (pMP->isBad()) { *vit = static_cast<MapPoint*>(NULL); continue; }
    pMP->IncreaseVisible();
    pMP->mnLastFrameSeen = mCurrentFrame.mnId;
    pMP->mbTrackInView = false;
    pMP->mbTrackInViewR = false;
  }
  // candidates of the projection test (:3368-3376)
  std::vector<MapPoint*> cand;
  cand.reserve(mvpLocalMapPoints.size());
  for (MapPoint* pMP : mvpLocalMapPoints) {
    if (pMP->mnLastFrameSeen == mCurrentFrame.mnId) continue;
    if (pMP->isBad()) continue;
    cand.push_back(pMP);
  }
  const int n = (int)cand.size();
  std::vector<float> pos(3 * n), nrm(3 * n), dmin(n), dmax(n);
  for (int i = 0; i < n; i++) {
    const Eigen::Vector3f P = cand[i]->GetWorldPos(), N = cand[i]->GetNormal();
    for (int k = 0; k < 3; k++) { pos[3 * i + k] = P[k]; nrm[3 * i + k] = N[k]; }
    // the view carries the raw mfMinDistance / mfMaxDistance (PredictScale divides the raw maximum by the
    // distance, MapPoint.cc:531-546): two one-line public getters added to include/MapPoint.h,
    // because dividing Get{Min,Max}DistanceInvariance() by 0.8f / 1.2f does not round-trip in float
    dmin[i] = cand[i]->GetMinDistanceRaw();
    dmax[i] = cand[i]->GetMaxDistanceRaw();
  }
  orb_frustum_view v;
  v.n = n;
  v.world_pos = pos.data(); v.normal = nrm.data(); v.min_dist = dmin.data(); v.max_dist = dmax.data();
  // mRcw / mtcw are private; Frame::UpdatePoseMatrices (Frame.cc:472-479) sets them to exactly these two expressions
  const Sophus::SE3<float> Tcw = mCurrentFrame.GetPose();
  const Eigen::Matrix3f R = Tcw.rotationMatrix();
  const Eigen::Vector3f t = Tcw.translation(), Ow = mCurrentFrame.GetOw();
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) v.Rcw[3 * r + c] = R(r, c);
  for (int k = 0; k < 3; k++) { v.tcw[k] = t[k]; v.Ow[k] = Ow[k]; }
  v.fx = mCurrentFrame.fx; v.fy = mCurrentFrame.fy; v.cx = mCurrentFrame.cx; v.cy = mCurrentFrame.cy;
  v.bf = mCurrentFrame.mbf;
  v.min_x = Frame::mnMinX; v.max_x = Frame::mnMaxX; v.min_y = Frame::mnMinY; v.max_y = Frame::mnMaxY;
  v.log_scale_factor = mCurrentFrame.mfLogScaleFactor;
  v.n_levels = mCurrentFrame.mnScaleLevels;

  static thread_local orb_frustum* h = nullptr;
  if (!h && frustum_create(/*device=*/0, &h) != ORB_OK)
    throw std::runtime_error(std::string("frustum_create: ") + orb_last_error());
  // start from the members as they are, so that what the reference leaves stale stays stale
  std::vector<uint8_t> in_view(n);
  std::vector<float> px(n), py(n), pxr(n), vcos(n), depth(n);
  std::vector<int32_t> lvl(n);
  for (int i = 0; i < n; i++) {
    pxr[i] = cand[i]->mTrackProjXR; vcos[i] = cand[i]->mTrackViewCos; depth[i] = cand[i]->mTrackDepth;
    lvl[i] = cand[i]->mnTrackScaleLevel;
  }
  int nToMatch = frame_is_in_frustum(h, &v, 0.5f, in_view.data(), px.data(), py.data(), pxr.data(), lvl.data(),
                                     vcos.data(), depth.data());
  if (nToMatch < 0) throw std::runtime_error(std::string("frame_is_in_frustum: ") + orb_last_error());
  for (int i = 0; i < n; i++) {
    MapPoint* pMP = cand[i];
    pMP->mbTrackInView = in_view[i] != 0;
    pMP->mTrackProjX = px[i]; pMP->mTrackProjY = py[i];
    pMP->mTrackProjXR = pxr[i]; pMP->mTrackViewCos = vcos[i]; pMP->mTrackDepth = depth[i];
    pMP->mnTrackScaleLevel = lvl[i];
    if (in_view[i]) pMP->IncreaseVisible();                 // :3380
    if (pMP->mbTrackInView) mCurrentFrame.mmProjectPoints[pMP->mnId] = cv::Point2f(pMP->mTrackProjX, pMP->mTrackProjY);  // :3383-3386
  }
  if (nToMatch > 0) {                                        // :3389-3422, unchanged
    ORBmatcher matcher(0.8);
    int th = 1;
    if (mSensor == System::RGBD || mSensor == System::IMU_RGBD) th = 3;
    if (mpAtlas->isImuInitialized()) {
      if (mpAtlas->GetCurrentMap()->GetIniertialBA2()) th = 2;
      else th = 6;
    } else if (!mpAtlas->isImuInitialized() && (mSensor == System::IMU_MONOCULAR || mSensor == System::IMU_STEREO ||
                                                mSensor == System::IMU_RGBD)) {
      th = 10;
    }
    if (mCurrentFrame.mnId < mnLastRelocFrameId + 2) th = 5;  // relocalised recently: coarser search
    if (mState == LOST || mState == RECENTLY_LOST) th = 15;
    matcher.SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th, mpLocalMapper->mbFarPoints,
                               mpLocalMapper->mThFarPoints);
  }
}

}  // namespace ORB_SLAM3
