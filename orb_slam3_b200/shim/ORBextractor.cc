// Drop-in replacement for the reference's src/ORBextractor.cc: the same class
// (include/ORBextractor.h:49-107, compiled unchanged) whose methods forward to
// the C ABI of liborbb200.so.  Build it inside the ORB_SLAM3 tree *instead of*
// src/ORBextractor.cc and link with -lorbb200.  Frame.cc,
// Tracking.cc and every other caller stay untouched.
//
//   ORBextractor::ORBextractor   <- ORBextractor.cc:409-469 (tables come from orb_create's host code)
//   ORBextractor::operator()     <- ORBextractor.cc:1086-1168 (orb_extract)
//   mvImagePyramid               <- refreshed from orb_pyramid() after every call (ORBextractor.h:83;
//                                   Frame::ComputeStereoMatches reads it, Frame.cc:818-923)
#include "ORBextractor.h"

#include <mutex>
#include <stdexcept>
#include <unordered_map>

#include "orb_b200.h"

namespace ORB_SLAM3 {

namespace {
// The reference class has no spare member for the engine handle and its header
// must stay byte-identical, so handles live in a side table keyed by `this`.
std::mutex g_mu;
std::unordered_map<const ORBextractor*, orb_extractor*> g_handles;

orb_extractor* handle_of(const ORBextractor* self) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(self);
  return it == g_handles.end() ? nullptr : it->second;
}
}  // namespace

// Used by shim/Frame_ComputeStereoMatches.cc: the engine handle behind a reference extractor object.
orb_extractor* orbb200_handle_of(const ORBextractor* self) { return handle_of(self); }

// The reference header declares an empty inline destructor (ORBextractor.h:52), so the engine cannot be freed
// from ~ORBextractor without touching the header: Tracking::~Tracking / System::Shutdown call this hook (one
// line each); an extractor constructed at a recycled address releases a stale entry itself.
void orbb200_release(const ORBextractor* self) {
  orb_extractor* h = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(self);
    if (it == g_handles.end()) return;
    h = it->second;
    g_handles.erase(it);
  }
  orb_destroy(h);
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
      minThFAST(_minThFAST) {
  orb_extractor* h = nullptr;
  if (orb_create(_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, /*device=*/0, &h) != ORB_OK)
    throw std::runtime_error(std::string("orb_create: ") + orb_last_error());
  mvScaleFactor.resize(nlevels);
  mvInvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);
  orb_get_scale_factors(h, mvScaleFactor.data());
  orb_get_inverse_scale_factors(h, mvInvScaleFactor.data());
  orb_get_scale_sigma_squares(h, mvLevelSigma2.data());
  orb_get_inverse_scale_sigma_squares(h, mvInvLevelSigma2.data());
  orb_get_features_per_level(h, mnFeaturesPerLevel.data());
  mvImagePyramid.resize(nlevels);
  orbb200_release(this);  // an earlier extractor lived at this address and was never released
  std::lock_guard<std::mutex> lk(g_mu);
  g_handles[this] = h;
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
  if (_image.empty()) return -1;  // ORBextractor.cc:1090-1091
  cv::Mat image = _image.getMat();
  assert(image.type() == CV_8UC1);
  orb_extractor* h = handle_of(this);
  const int cap = nfeatures + 8 * nlevels + 64;  // quota + <=3 overshoot per level
  // cv::KeyPoint and orb_keypoint share one 28-byte layout
  static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
  std::vector<cv::KeyPoint> kps(cap);
  cv::Mat desc(cap, 32, CV_8U);
  int n = 0;
  const int mono = orb_extract(h, image.data, image.rows, image.cols, image.step, vLappingArea[0], vLappingArea[1],
                               reinterpret_cast<orb_keypoint*>(kps.data()), desc.data, cap, &n);
  if (mono < 0) throw std::runtime_error(std::string("orb_extract: ") + orb_last_error());
  kps.resize(n);
  _keypoints.swap(kps);
  if (n == 0) {
    _descriptors.release();  // :1108-1109
  } else {
    _descriptors.create(n, 32, CV_8U);
    desc.rowRange(0, n).copyTo(_descriptors.getMat());
  }
  // mvImagePyramid (ORBextractor.h:83): the only reader in the reference is Frame::ComputeStereoMatches
  // (Frame.cc:818-923).  With shim/Frame_ComputeStereoMatches.cc built in (ORB_B200_STEREO_ON_DEVICE) nothing
  // reads the host mirror and the full-pyramid D2H per frame is skipped; otherwise the levels are refreshed as
  // unpadded views into engine-owned pinned memory, valid until the next call on this extractor.
#ifndef ORB_B200_STEREO_ON_DEVICE
  for (int level = 0; level < nlevels; ++level) {
    const uint8_t* p = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    if (orb_pyramid(h, 0, level, &p, &rows, &cols, &step) == ORB_OK)
      mvImagePyramid[level] = cv::Mat(rows, cols, CV_8UC1, const_cast<uint8_t*>(p), step);
  }
#endif
  return mono;
}

// The remaining protected members of the reference class are only called from
// operator() in the reference; they are kept as no-ops so the header links.
void ORBextractor::ComputePyramid(cv::Mat) {}
void ORBextractor::ComputeKeyPointsOctTree(std::vector<std::vector<cv::KeyPoint> >&) {}
void ORBextractor::ComputeKeyPointsOld(std::vector<std::vector<cv::KeyPoint> >&) {}
std::vector<cv::KeyPoint> ORBextractor::DistributeOctTree(const std::vector<cv::KeyPoint>&, const int&, const int&,
                                                          const int&, const int&, const int&, const int&) {
  return std::vector<cv::KeyPoint>();
}
void ExtractorNode::DivideNode(ExtractorNode&, ExtractorNode&, ExtractorNode&, ExtractorNode&) {}

}  // namespace ORB_SLAM3
