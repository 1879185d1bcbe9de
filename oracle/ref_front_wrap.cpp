// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's front-end object code (oracle/_ref/libref_front.so):
// /root/reference/src/ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc, ORBextractor.cc and CameraModels/Pinhole.cpp
// compiled UNMODIFIED against the stand-ins of oracle/cvcompat/ (functional 8-bit cv::Mat) and oracle/eigencompat/
// (functional fixed-size Eigen, Sophus SE3 with so3.hpp / se3.hpp's formulas, g2o members); everything else those
// translation units name is bound to stubs (oracle/Makefile) that print the symbol and abort if a run ever reaches them.
// Real Frame / KeyFrame / MapPoint objects are filled from the flat views of include/orb_b200.h and the reference's own
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)   (ORBmatcher.cc:40-285)
//   ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)                        (:1669-1890)
//   ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, pairs, bOnlyStereo, bCoarse)           (:907-1146)
//   Frame::isInFrustum(MapPoint*, viewingCosLimit) + MapPoint::PredictScale                         (Frame.cc:542-612)
//   Frame::ComputeStereoMatches()                                                                   (Frame.cc:811-981)
//   Frame::GetFeaturesInArea / AssignFeaturesToGrid, Pinhole::epipolarConstrain
// run as object code.  tests/test_ref_front.py holds the oracle (orc_match.cpp, orc_frustum.cpp, orc_stereo.cpp) against
// them.  Nothing in the product links this.
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

// the reference keeps the tracking state of a MapPoint and the pose of a Frame private; this translation unit fills them
// directly (access specifiers do not change the object layout the unmodified translation units were compiled with)
#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#include "Pinhole.h"
#undef private
#undef protected

#include "../include/orb_b200.h"

using namespace ORB_SLAM3;

namespace {

cv::Mat desc_rows(const uint8_t* d, int n) {
  cv::Mat m(std::max(n, 1), 32, CV_8U);
  if (n > 0) memcpy(m.ptr(), d, 32 * (size_t)n);
  return n > 0 ? m : m.rowRange(0, 0);
}

struct World {  // owns what a call allocates
  std::vector<MapPoint*> points;
  std::vector<GeometricCamera*> cams;
  MapPoint* taken = nullptr;
  ~World() { for (MapPoint* p : points) delete p; for (GeometricCamera* c : cams) delete c; }
  MapPoint* point() { points.push_back(new MapPoint()); return points.back(); }
};

void fill_frame(Frame& F, const orb_frame_view* v, World& W) {
  F.N = v->n;
  F.mvKeysUn.resize(v->n);
  static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "KeyPoint layout");
  if (v->n) memcpy((void*)F.mvKeysUn.data(), v->keys, sizeof(orb_keypoint) * (size_t)v->n);
  F.mvKeys = F.mvKeysUn;
  F.mvuRight.assign(v->n, -1.f);
  if (v->u_right) F.mvuRight.assign(v->u_right, v->u_right + v->n);
  F.mvDepth.assign(v->n, -1.f);
  F.mDescriptors = desc_rows(v->desc, v->n);
  F.mvbOutlier.assign(v->n, false);
  F.mvpMapPoints.assign(v->n, static_cast<MapPoint*>(NULL));
  if (v->kp_taken) {
    W.taken = W.point();
    W.taken->nObs = 1;  // mvpMapPoints[i] && Observations() > 0 (ORBmatcher.cc:88-90, :1747-1749)
    for (int i = 0; i < v->n; i++) if (v->kp_taken[i]) F.mvpMapPoints[i] = W.taken;
  }
  F.Nleft = -1; F.Nright = -1;
  F.mnScaleLevels = v->n_levels;
  F.mvScaleFactors.assign(v->scale_factors, v->scale_factors + v->n_levels);
  F.mvLevelSigma2.assign(v->level_sigma2, v->level_sigma2 + v->n_levels);
  F.mvInvScaleFactors.resize(v->n_levels); F.mvInvLevelSigma2.resize(v->n_levels);
  for (int l = 0; l < v->n_levels; l++) { F.mvInvScaleFactors[l] = 1.0f / F.mvScaleFactors[l]; F.mvInvLevelSigma2[l] = 1.0f / F.mvLevelSigma2[l]; }
  F.mfScaleFactor = v->n_levels > 1 ? v->scale_factors[1] : 1.2f;
  F.mfLogScaleFactor = log(F.mfScaleFactor);
  Frame::mnMinX = v->min_x; Frame::mnMinY = v->min_y; Frame::mnMaxX = v->max_x; Frame::mnMaxY = v->max_y;
  Frame::mfGridElementWidthInv = v->grid_w_inv; Frame::mfGridElementHeightInv = v->grid_h_inv;
  Frame::fx = v->fx; Frame::fy = v->fy; Frame::cx = v->cx; Frame::cy = v->cy;
  Frame::invfx = 1.0f / v->fx; Frame::invfy = 1.0f / v->fy;
  F.mbf = v->bf; F.mb = v->b;
  W.cams.push_back(new Pinhole(std::vector<float>{v->fx, v->fy, v->cx, v->cy}));
  F.mpCamera = W.cams.back();
  F.mpCamera2 = nullptr;
  F.AssignFeaturesToGrid();
}

}  // namespace

extern "C" {

// ORBmatcher(nn_ratio).SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints); assign_out as match_project_local
int ref_front_project_local(const orb_frame_view* fv, const orb_mappoint_view* mv, float th, float nn_ratio, int far_points,
                            float th_far, int32_t* assign_out) {
  World W;
  Frame F;
  fill_frame(F, fv, W);
  std::vector<MapPoint*> mps(mv->n);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < mv->n; i++) {
    MapPoint* p = W.point();
    p->mbTrackInView = mv->track_in_view[i]; p->mbTrackInViewR = false;
    p->mbBad = mv->is_bad[i]; p->nObs = mv->has_obs[i] ? 1 : 0;
    p->mTrackProjX = mv->proj_x[i]; p->mTrackProjY = mv->proj_y[i]; p->mTrackProjXR = mv->proj_xr[i];
    p->mnTrackScaleLevel = mv->scale_level[i]; p->mTrackViewCos = mv->view_cos[i]; p->mTrackDepth = mv->depth[i];
    p->mDescriptor = desc_rows(mv->desc + 32 * (size_t)i, 1);
    mps[i] = p; index[p] = i;
  }
  ORBmatcher matcher(nn_ratio);
  const int n = matcher.SearchByProjection(F, mps, th, far_points != 0, th_far);
  for (int i = 0; i < fv->n; i++) {
    auto it = index.find(F.mvpMapPoints[i]);
    assign_out[i] = it == index.end() ? -1 : it->second;
  }
  return n;
}

// ORBmatcher(0.9, check_orientation).SearchByProjection(Cur, Last, th, bMono); LastFrame's pose is chosen so that the
// reference's own bForward / bBackward tests (:1692-1693) come out as the caller says.  assign_out as match_project_last.
int ref_front_project_last(const orb_frame_view* cur, const orb_lastframe_view* last, const float* Tcw_qt7, int forward,
                           int backward, float th, int check_orientation, float nn_ratio, int32_t* assign_out) {
  World W;
  Frame C, L;
  fill_frame(C, cur, W);
  const Sophus::SE3f Tcw(Eigen::Quaternionf(Tcw_qt7[3], Tcw_qt7[0], Tcw_qt7[1], Tcw_qt7[2]), Eigen::Vector3f(Tcw_qt7[4], Tcw_qt7[5], Tcw_qt7[6]));
  C.SetPose(Tcw);
  L.N = last->n;
  L.Nleft = -1; L.Nright = -1;
  L.mvKeys.resize(last->n); L.mvKeysUn.resize(last->n);
  L.mvbOutlier.assign(last->n, false);
  L.mvpMapPoints.assign(last->n, static_cast<MapPoint*>(NULL));
  std::map<MapPoint*, int> index;
  for (int i = 0; i < last->n; i++) {
    L.mvKeys[i].octave = last->octave[i]; L.mvKeysUn[i].octave = last->octave[i];
    L.mvKeys[i].angle = last->angle[i]; L.mvKeysUn[i].angle = last->angle[i];
    if (!last->has_mp[i]) continue;
    MapPoint* p = W.point();
    p->nObs = last->has_obs[i] ? 1 : 0;
    p->mWorldPos = Eigen::Vector3f(last->world_pos[3 * i], last->world_pos[3 * i + 1], last->world_pos[3 * i + 2]);
    p->mDescriptor = desc_rows(last->desc + 32 * (size_t)i, 1);
    L.mvpMapPoints[i] = p; index[p] = i;
  }
  const Eigen::Vector3f twc = Tcw.inverse().translation();
  const float want = forward ? 2.f * C.mb + 1.f : (backward ? -(2.f * C.mb + 1.f) : 0.f);  // tlc(2) of :1690
  L.SetPose(Sophus::SE3f(Eigen::Quaternionf(1.f, 0.f, 0.f, 0.f), Eigen::Vector3f(-twc[0], -twc[1], want - twc[2])));
  ORBmatcher matcher(nn_ratio, check_orientation != 0);
  const int n = matcher.SearchByProjection(C, L, th, !(forward || backward));
  for (int i = 0; i < cur->n; i++) {
    auto it = index.find(C.mvpMapPoints[i]);   // NULL both for "never matched" and "cleared by the rotation check" (:1875-1884)
    assign_out[i] = it == index.end() ? -1 : it->second;
  }
  return n;
}

// Frame::isInFrustum for every map point of the view; outputs as frame_is_in_frustum
int ref_front_is_in_frustum(const orb_frustum_view* v, float viewing_cos_limit, uint8_t* track_in_view, float* proj_x, float* proj_y,
                            float* proj_xr, int32_t* scale_level, float* view_cos, float* depth) {
  World W;
  Frame F;
  F.N = 0; F.Nleft = -1; F.Nright = -1;
  F.mnScaleLevels = v->n_levels; F.mfLogScaleFactor = v->log_scale_factor;
  Frame::mnMinX = v->min_x; Frame::mnMaxX = v->max_x; Frame::mnMinY = v->min_y; Frame::mnMaxY = v->max_y;
  Frame::fx = v->fx; Frame::fy = v->fy; Frame::cx = v->cx; Frame::cy = v->cy;
  F.mbf = v->bf;
  W.cams.push_back(new Pinhole(std::vector<float>{v->fx, v->fy, v->cx, v->cy}));
  F.mpCamera = W.cams.back(); F.mpCamera2 = nullptr;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) F.mRcw(r, c) = v->Rcw[3 * r + c]; F.mtcw[r] = v->tcw[r]; F.mOw[r] = v->Ow[r]; }
  int n_in = 0;
  for (int i = 0; i < v->n; i++) {
    MapPoint* p = W.point();
    p->mWorldPos = Eigen::Vector3f(v->world_pos[3 * i], v->world_pos[3 * i + 1], v->world_pos[3 * i + 2]);
    p->mNormalVector = Eigen::Vector3f(v->normal[3 * i], v->normal[3 * i + 1], v->normal[3 * i + 2]);
    p->mfMinDistance = v->min_dist[i]; p->mfMaxDistance = v->max_dist[i];
    const bool in = F.isInFrustum(p, viewing_cos_limit);
    track_in_view[i] = p->mbTrackInView; proj_x[i] = p->mTrackProjX; proj_y[i] = p->mTrackProjY;
    if (in) { proj_xr[i] = p->mTrackProjXR; scale_level[i] = p->mnTrackScaleLevel; view_cos[i] = p->mTrackViewCos; depth[i] = p->mTrackDepth; n_in++; }
  }
  return n_in;
}

// Frame::ComputeStereoMatches (Frame.cc:811-981) on a Frame holding the caller's keypoints / descriptors and two ORBextractor
// objects whose mvImagePyramid are headers over the caller's level images.  Returns the number of keypoints left with a
// stereo match; u_right / depth = mvuRight / mvDepth.
int ref_front_stereo_match(int nl, const orb_keypoint* kl, const uint8_t* dl, int nr, const orb_keypoint* kr, const uint8_t* dr,
                           int n_levels, const uint8_t* const* pyr_l, const uint8_t* const* pyr_r, const int32_t* lw,
                           const int32_t* lh, const int32_t* ls, float scale_factor, float bf, float b, float* u_right, float* depth) {
  ORBextractor exl(1000, scale_factor, n_levels, 20, 7), exr(1000, scale_factor, n_levels, 20, 7);
  for (int l = 0; l < n_levels; l++) {
    exl.mvImagePyramid[l] = cv::Mat(lh[l], lw[l], CV_8UC1, (void*)pyr_l[l], (size_t)ls[l]);
    exr.mvImagePyramid[l] = cv::Mat(lh[l], lw[l], CV_8UC1, (void*)pyr_r[l], (size_t)ls[l]);
  }
  Frame F;
  F.N = nl; F.Nleft = -1; F.Nright = -1;
  F.mpORBextractorLeft = &exl; F.mpORBextractorRight = &exr;
  F.mvKeys.resize(nl); F.mvKeysRight.resize(nr);
  if (nl) memcpy((void*)F.mvKeys.data(), kl, sizeof(orb_keypoint) * (size_t)nl);
  if (nr) memcpy((void*)F.mvKeysRight.data(), kr, sizeof(orb_keypoint) * (size_t)nr);
  F.mDescriptors = desc_rows(dl, nl); F.mDescriptorsRight = desc_rows(dr, nr);
  F.mvScaleFactors = exl.GetScaleFactors(); F.mvInvScaleFactors = exl.GetInverseScaleFactors();
  F.mbf = bf; F.mb = b;
  F.ComputeStereoMatches();
  int kept = 0;
  for (int i = 0; i < nl; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; kept += F.mvDepth[i] > 0; }
  F.mpORBextractorLeft = nullptr; F.mpORBextractorRight = nullptr;
  return kept;
}

}  // extern "C"

namespace {
template <typename T> T& writable(const T& x) { return const_cast<T&>(x); }   // KeyFrame's descriptive members are const

void fill_keyframe(KeyFrame& K, const orb_frame_view* v, const orb_featvec_view* fv, const float* Tcw_qt7, World& W) {
  writable(K.N) = v->n; writable(K.NLeft) = -1; writable(K.NRight) = -1;
  std::vector<cv::KeyPoint>& keys = writable(K.mvKeysUn);
  keys.resize(v->n);
  if (v->n) memcpy((void*)keys.data(), v->keys, sizeof(orb_keypoint) * (size_t)v->n);
  writable(K.mvKeys) = keys;
  std::vector<float>& ur = writable(K.mvuRight);
  ur.assign(v->n, -1.f);
  if (v->u_right) ur.assign(v->u_right, v->u_right + v->n);
  writable(K.mDescriptors) = desc_rows(v->desc, v->n);
  writable(K.mvScaleFactors).assign(v->scale_factors, v->scale_factors + v->n_levels);
  writable(K.mvLevelSigma2).assign(v->level_sigma2, v->level_sigma2 + v->n_levels);
  K.mvpMapPoints.assign(v->n, static_cast<MapPoint*>(NULL));
  if (v->kp_taken) {
    if (!W.taken) W.taken = W.point();
    for (int i = 0; i < v->n; i++) if (v->kp_taken[i]) K.mvpMapPoints[i] = W.taken;   // "there is already a MapPoint" (:972-976)
  }
  for (int k = 0; k < fv->n_nodes; k++) {
    std::vector<unsigned int>& f = K.mFeatVec[fv->node_ids[k]];
    for (int p = fv->ptr[k]; p < fv->ptr[k + 1]; p++) f.push_back((unsigned int)fv->idx[p]);
  }
  W.cams.push_back(new Pinhole(std::vector<float>{v->fx, v->fy, v->cx, v->cy}));
  K.mpCamera = W.cams.back(); K.mpCamera2 = nullptr;
  K.SetPose(Sophus::SE3f(Eigen::Quaternionf(Tcw_qt7[3], Tcw_qt7[0], Tcw_qt7[1], Tcw_qt7[2]), Eigen::Vector3f(Tcw_qt7[4], Tcw_qt7[5], Tcw_qt7[6])));
}
}  // namespace

extern "C" {

// ORBmatcher(nn_ratio, check_orientation).SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) on two real
// KeyFrame objects (Pinhole, no second camera).  pairs_out = vMatchedPairs; F12_out (row-major) and ep_out are the matrix
// Pinhole::epipolarConstrain builds from (R12, t12) (Pinhole.cpp:107-112) and the epipole of :917-920 -- the inputs the C ABI
// takes from its caller -- evaluated by the same expressions.
int ref_front_triangulate(const orb_frame_view* kf1, const orb_frame_view* kf2, const orb_featvec_view* fv1, const orb_featvec_view* fv2,
                          const float* T1w_qt7, const float* T2w_qt7, int only_stereo, int coarse, int check_orientation,
                          int32_t* pairs_out, int cap, float* F12_out, float* ep_out) {
  World W;
  KeyFrame K1, K2;
  fill_keyframe(K1, kf1, fv1, T1w_qt7, W);
  fill_keyframe(K2, kf2, fv2, T2w_qt7, W);
  {
    const Sophus::SE3f T12 = K1.GetPose() * K2.GetPoseInverse();
    const Eigen::Matrix3f R12 = T12.rotationMatrix();
    const Eigen::Vector3f t12 = T12.translation();
    const Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
    const Eigen::Matrix3f Ka = K1.mpCamera->toK_();
    const Eigen::Matrix3f Kb = K2.mpCamera->toK_();
    const Eigen::Matrix3f F12 = Ka.transpose().inverse() * t12x * R12 * Kb.inverse();
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F12_out[3 * r + c] = F12(r, c);
    const Eigen::Vector3f C2 = K2.GetPose() * K1.GetCameraCenter();
    const Eigen::Vector2f ep = K2.mpCamera->project(C2);
    ep_out[0] = ep(0); ep_out[1] = ep(1);
  }
  ORBmatcher matcher(0.6f, check_orientation != 0);
  std::vector<std::pair<size_t, size_t> > pairs;
  const int n = matcher.SearchForTriangulation(&K1, &K2, pairs, only_stereo != 0, coarse != 0);
  for (size_t i = 0; i < pairs.size() && (int)i < cap; i++) { pairs_out[2 * i] = (int32_t)pairs[i].first; pairs_out[2 * i + 1] = (int32_t)pairs[i].second; }
  return n;
}

}  // extern "C"
