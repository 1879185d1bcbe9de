// TEST INFRASTRUCTURE ONLY (see orc_common.h) -- CPU restatement of
// Frame::ComputeStereoMatches (reference src/Frame.cc:811-981), SURVEY.md 8(f-1),
// on flat arrays: the two keypoint / descriptor sets an ORBextractor pair
// produced (mvKeys / mvKeysRight, not undistorted) and the two un-blurred image
// pyramids (mvImagePyramid).  PARITY UNPINNED BY THE REFERENCE (no tests ship
// with it); cv::norm(NORM_L1) and std::sort are restated with plain loops and
// checked against cv2 / numpy in tests/test_stereo_oracle.py.
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "orc_common.h"

namespace {

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074)
int hamming256(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4);
    memcpy(&pb, b + 4 * i, 4);
    dist += __builtin_popcount(pa ^ pb);
  }
  return dist;
}

const int TH_HIGH = 100, TH_LOW = 50;  // ORBmatcher.cc:35-36

}  // namespace

extern "C" {

// u_right / depth: n_left floats (mvuRight / mvDepth); sad (optional): the L1
// window distance of every left keypoint that reached the final list before the
// median test (vDistIdx), -1 otherwise.  Returns the number of stereo matches kept.
int orc_stereo_match(int n_left, const orc_keypoint* kl, const uint8_t* dl, int n_right, const orc_keypoint* kr,
                     const uint8_t* dr, int nlevels, const uint8_t* const* pyr_l, const uint8_t* const* pyr_r,
                     const int* lw, const int* lh, const int* lstep, const float* scale, const float* inv_scale,
                     float bf, float b, float* u_right, float* depth, int* sad) {
  (void)nlevels;
  const int N = n_left;
  for (int i = 0; i < N; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; if (sad) sad[i] = -1; }  // :813-814
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                                               // :816
  const int nRows = lh[0];                                                                    // :818

  // row table (:820-838); the reference indexes vRowIndices without a bounds
  // check -- keypoints sit >= 16 level pixels from the border and r <= 2*scale,
  // so the clamp below never triggers on extractor output
  std::vector<std::vector<int>> rows(nRows);
  for (int iR = 0; iR < n_right; iR++) {
    const float kpY = kr[iR].y;
    const float r = 2.0f * scale[kr[iR].octave];
    const int maxr = (int)ceilf(kpY + r);
    const int minr = (int)floorf(kpY - r);
    for (int yi = std::max(minr, 0); yi <= std::min(maxr, nRows - 1); yi++) rows[yi].push_back(iR);
  }

  const float minZ = b, minD = 0, maxD = bf / minZ;  // :841-843
  std::vector<std::pair<int, int>> vDistIdx;
  vDistIdx.reserve(N);

  for (int iL = 0; iL < N; iL++) {
    const int levelL = kl[iL].octave;
    const float vL = kl[iL].y, uL = kl[iL].x;
    const int row = std::min(std::max((int)vL, 0), nRows - 1);
    const std::vector<int>& cand = rows[row];  // :856
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;
    int bestIdxR = 0;
    const uint8_t* d1 = dl + (size_t)iL * 32;
    for (size_t iC = 0; iC < cand.size(); iC++) {  // :873-895
      const int iR = cand[iC];
      if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
      const float uR = kr[iR].x;
      if (uR >= minU && uR <= maxU) {
        const int dist = hamming256(d1, dr + (size_t)iR * 32);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (!(bestDist < thOrbDist)) continue;  // :898

    // sub-pixel refinement by correlation at the keypoint's pyramid level (:900-947)
    const float uR0 = kr[bestIdxR].x;
    const float scaleFactor = inv_scale[levelL];
    const float scaleduL = roundf(uL * scaleFactor);
    const float scaledvL = roundf(vL * scaleFactor);
    const float scaleduR0 = roundf(uR0 * scaleFactor);
    const int w = 5, L = 5;
    const uint8_t* IL = pyr_l[levelL];
    const uint8_t* IR = pyr_r[levelL];
    const int step = lstep[levelL], cols = lw[levelL];
    int bestSad = INT_MAX, bestincR = 0;
    float vDists[2 * 5 + 1];
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= cols) continue;  // :918-919
    const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
    for (int incR = -L; incR <= +L; incR++) {
      const int xr0 = (int)(scaleduR0 + incR - w);
      int s = 0;  // cv::norm(IL, IR, NORM_L1) on two 11x11 CV_8U windows
      for (int yy = 0; yy < 2 * w + 1; yy++)
        for (int xx = 0; xx < 2 * w + 1; xx++)
          s += abs((int)IL[(size_t)(y0 + yy) * step + xl0 + xx] - (int)IR[(size_t)(y0 + yy) * step + xr0 + xx]);
      const float dist = (float)s;
      if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
      vDists[L + incR] = dist;
    }
    if (bestincR == -L || bestincR == L) continue;  // :935-936
    const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);  // :950
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }  // double literals as in :956-960
      depth[iL] = bf / disparity;
      u_right[iL] = bestuR;
      vDistIdx.push_back(std::pair<int, int>(bestSad, iL));
    }
  }

  // median-based rejection (:968-982); the reference reads vDistIdx[size/2] even
  // when the list is empty (undefined) -- here an empty list rejects nothing
  if (vDistIdx.empty()) return 0;
  std::sort(vDistIdx.begin(), vDistIdx.end());
  if (sad) for (auto& p : vDistIdx) sad[p.second] = p.first;
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)vDistIdx.size();
  for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
    if (vDistIdx[i].first < thDist) break;
    u_right[vDistIdx[i].second] = -1;
    depth[vDistIdx[i].second] = -1;
    kept--;
  }
  return kept;
}

}  // extern "C"
