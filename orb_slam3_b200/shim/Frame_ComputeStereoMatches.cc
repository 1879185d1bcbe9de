// Drop-in replacement for Frame::ComputeStereoMatches (reference src/Frame.cc:811-981,
// SURVEY.md 8(f-1)).  Compile it inside the ORB_SLAM3 tree next to shim/ORBextractor.cc and
// wrap the reference body in `#ifndef ORB_B200_HOTPATH ... #endif`.
//
// The reference walks mvKeys / mvKeysRight / the two descriptor matrices and both
// extractors' mvImagePyramid on the host.  All of that is still resident on the
// device after the two ORBextractor::operator() calls of the Frame constructor
// (Frame.cc:122-125), so the shim only passes the two engine handles plus mbf / mb
// and receives mvuRight / mvDepth; with this unit in place nothing reads the host
// pyramid mirror any more.
#include <stdexcept>
#include <string>

#include "Frame.h"
#include "orb_b200.h"

namespace ORB_SLAM3 {

orb_extractor* orbb200_handle_of(const ORBextractor* self);  // shim/ORBextractor.cc

void Frame::ComputeStereoMatches() {
  mvuRight = std::vector<float>(N, -1.0f);  // :813-814
  mvDepth = std::vector<float>(N, -1.0f);
  if (N == 0) return;
  // one handle per calling thread: stereo frames are only built by the Tracking thread
  static thread_local orb_stereo* st = nullptr;
  if (!st && stereo_create(/*device=*/0, &st) != ORB_OK)
    throw std::runtime_error(std::string("stereo_create: ") + orb_last_error());
  orb_extractor* hl = orbb200_handle_of(mpORBextractorLeft);
  orb_extractor* hr = orbb200_handle_of(mpORBextractorRight);
  // keypoints are matched in the extractor's output order, which is the order of mvKeys
  const int rc = stereo_match(st, hl, hr, mbf, mb, mvuRight.data(), mvDepth.data(), N);
  if (rc < 0) throw std::runtime_error(std::string("stereo_match: ") + orb_last_error());
}

}  // namespace ORB_SLAM3
