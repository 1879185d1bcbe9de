// TEST INFRASTRUCTURE ONLY (see orc_common.h).  CPU restatement of the
// reference's ORB front-end, used as the parity oracle and the CPU baseline.
//
// Restates (paths relative to /root/reference):
//   src/ORBextractor.cc:76-103    IC_Angle
//   src/ORBextractor.cc:107-146   computeOrbDescriptor
//   src/ORBextractor.cc:409-469   ORBextractor::ORBextractor (tables, quotas, umax)
//   src/ORBextractor.cc:480-536   ExtractorNode::DivideNode
//   src/ORBextractor.cc:538-553   compareNodes
//   src/ORBextractor.cc:555-779   DistributeOctTree
//   src/ORBextractor.cc:781-896   ComputeKeyPointsOctTree
//   src/ORBextractor.cc:1086-1168 operator()
//   src/ORBextractor.cc:1170-1195 ComputePyramid
// and the un-vendored OpenCV (find_package(OpenCV 4.4), CMakeLists.txt:33)
// primitives those lines call, whose arithmetic is restated from the published
// OpenCV algorithm and pinned against cv2 4.13 in
// tests/test_oracle_vs_cv2.py:
//   cv::resize(INTER_LINEAR, 8UC1), cv::FAST(TYPE_9_16, nonmax),
//   cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101), cv::fastAtan2, cvRound.
//
// Canonical float semantics: strict IEEE-754 single precision, no FMA
// contraction (compile with -ffp-contract=off), round-half-even for cvRound.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <list>
#include <thread>
#include <stdexcept>
#include <vector>

#include "orc_common.h"

namespace {

const int PATCH_SIZE = 31;
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;

static const int kPattern[1024] = {
#include "pattern_31.inc"
};

// cvRound: round half to even.
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }

struct Image {
  int w = 0, h = 0;
  std::vector<uint8_t> px;  // dense, step == w
  uint8_t* row(int y) { return px.data() + (size_t)y * w; }
  const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
};

// ---------------------------------------------------------------- cv::resize
// INTER_LINEAR on 8UC1, OpenCV fixed-point path: 11-bit
// coefficient pairs, horizontal pass in int, vertical pass
// (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2.
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh,
                      int dstep) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> alpha(2 * dw), beta(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    alpha[2 * dx] = (short)std::min(32767, std::max(-32768, cv_round((1.f - fx) * 2048.f)));
    alpha[2 * dx + 1] = (short)std::min(32767, std::max(-32768, cv_round(fx * 2048.f)));
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    beta[2 * dy] = (short)std::min(32767, std::max(-32768, cv_round((1.f - fy) * 2048.f)));
    beta[2 * dy + 1] = (short)std::min(32767, std::max(-32768, cv_round(fy * 2048.f)));
  }
  std::vector<int> hbuf0(dw), hbuf1(dw);
  int cached0 = -1, cached1 = -1;
  auto hpass = [&](int sy, std::vector<int>& out) {
    const uint8_t* S = src + (size_t)sy * sstep;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx];
      int sx1 = std::min(sx + 1, sw - 1);
      out[dx] = S[sx] * alpha[2 * dx] + S[sx1] * alpha[2 * dx + 1];
    }
  };
  for (int dy = 0; dy < dh; dy++) {
    int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
    int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
    // tiny two-row cache (rows advance monotonically)
    if (cached1 == sy0) { std::swap(hbuf0, hbuf1); std::swap(cached0, cached1); }
    if (cached0 != sy0) { hpass(sy0, hbuf0); cached0 = sy0; }
    if (sy1 == sy0) { hbuf1 = hbuf0; cached1 = sy1; }
    else if (cached1 != sy1) { hpass(sy1, hbuf1); cached1 = sy1; }
    const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
    uint8_t* D = dst + (size_t)dy * dstep;
    for (int x = 0; x < dw; x++)
      D[x] = (uint8_t)((((b0 * (hbuf0[x] >> 4)) >> 16) + ((b1 * (hbuf1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

// ------------------------------------------------------------------ cv::FAST
// FAST-9/16 with score = (largest t for which the pixel is still a corner) and
// strict 3x3 non-max suppression; only pixels at distance >= 3 from the border
// of the image passed in are tested; output row-major.
static const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1},
                                   {2, -2}, {1, -3},  {0, -3},  {-1, -3}, {-2, -2}, {-3, -1},
                                   {-3, 0}, {-3, 1},  {-2, 2},  {-1, 3}};

// Score of one pixel given its 16 ring differences d[k] = I(p) - I(ring k).
inline int fast_score_from_ring(const int* d /*16*/) {
  int best = 0;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) {
      int v = d[(k + j) & 15];
      mn = std::min(mn, v);
      mx = std::max(mx, v);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;  // corner at threshold t  <=>  best > t  <=>  score >= t
}

struct RawKp { int x, y; int score; };

void fast_detect(const uint8_t* img, int w, int h, int step, int threshold, bool nonmax,
                 std::vector<RawKp>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  const int bw = w - 6, bh = h - 6;  // tested band
  static thread_local std::vector<uint8_t> smap;  // score+1 clipped (score<=254), 0 = not a corner
  smap.assign((size_t)(bw + 2) * (bh + 2), 0);
  const int ss = bw + 2;
  int off[16];
  for (int k = 0; k < 16; k++) off[k] = kCircle[k][1] * step + kCircle[k][0];
  threshold = std::min(std::max(threshold, 0), 255);
  for (int y = 3; y < h - 3; y++) {
    const uint8_t* p = img + (size_t)y * step;
    uint8_t* srow = smap.data() + (size_t)(y - 3 + 1) * ss + 1;
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* c = p + x;
      const int v = c[0];
      const int hi = v + threshold, lo = v - threshold;
      // any 9-arc of the 16-ring covers >=2 of the 4 compass points and both
      // ends of at least one diameter-adjacent test; cheap rejection first
      int a = c[off[0]], b = c[off[8]];
      bool bright_possible = (a > hi) || (b > hi);
      bool dark_possible = (a < lo) || (b < lo);
      if (!bright_possible && !dark_possible) continue;
      int e = c[off[4]], f = c[off[12]];
      bright_possible = bright_possible && ((e > hi) || (f > hi));
      dark_possible = dark_possible && ((e < lo) || (f < lo));
      if (!bright_possible && !dark_possible) continue;
      uint32_t mb = 0, md = 0;
      int d[16];
      for (int k = 0; k < 16; k++) {
        int r = c[off[k]];
        d[k] = v - r;
        mb |= (uint32_t)(r > hi) << k;
        md |= (uint32_t)(r < lo) << k;
      }
      auto has_run9 = [](uint32_t m) {
        uint32_t m2 = m | (m << 16);
        uint32_t t = m2 & (m2 >> 1);
        t &= t >> 2;
        t &= t >> 4;
        t &= m2 >> 8;
        return t != 0;
      };
      if (!has_run9(mb) && !has_run9(md)) continue;
      int sc = fast_score_from_ring(d);
      srow[x - 3] = (uint8_t)(sc + 1 > 255 ? 255 : sc + 1);
    }
  }
  for (int y = 0; y < bh; y++) {
    const uint8_t* s0 = smap.data() + (size_t)(y) * ss + 1;
    const uint8_t* s1 = s0 + ss;
    const uint8_t* s2 = s1 + ss;
    for (int x = 0; x < bw; x++) {
      int s = s1[x];
      if (!s) continue;
      if (nonmax) {
        if (!(s > s1[x - 1] && s > s1[x + 1] && s > s0[x - 1] && s > s0[x] && s > s0[x + 1] &&
              s > s2[x - 1] && s > s2[x] && s > s2[x + 1]))
          continue;
      }
      out.push_back(RawKp{x + 3, y + 3, s - 1});
    }
  }
}

// -------------------------------------------------------- cv::GaussianBlur 7x7
// 8-bit fixed-point path: separable [18,34,48,56,48,34,18]/256,
// 8.8 horizontal then 8.8 vertical, one final rounding (v + 2^15) >> 16,
// BORDER_REFLECT_101.
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * n - 2 - p;
  }
  return p;
}

void gaussian_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep,
                    std::vector<uint16_t>& hb) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  hb.resize((size_t)w * h);  // caller-owned scratch: no per-call mmap/munmap
  std::vector<int> xi(w + 6);
  for (int x = -3; x < w + 3; x++) xi[x + 3] = reflect101(x, w);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src + (size_t)y * sstep;
    uint16_t* H = hb.data() + (size_t)y * w;
    for (int x = 0; x < w; x++) {
      int acc = 0;
      for (int k = 0; k < 7; k++) acc += K[k] * S[xi[x + k]];
      H[x] = (uint16_t)acc;
    }
  }
  for (int y = 0; y < h; y++) {
    const uint16_t* R[7];
    for (int k = 0; k < 7; k++) R[k] = hb.data() + (size_t)reflect101(y + k - 3, h) * w;
    uint8_t* D = dst + (size_t)y * dstep;
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int k = 0; k < 7; k++) acc += (uint32_t)K[k] * R[k][x];
      D[x] = (uint8_t)((acc + (1u << 15)) >> 16);
    }
  }
}

// -------------------------------------------------------------- cv::fastAtan2
// degree-scaled odd polynomial, every operation rounded to
// float, no FMA.
float fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / M_PI);
  const float p1 = 0.9997878412794807f * scale;
  const float p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale;
  const float p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ------------------------------------------------------------- the extractor
struct OctNode {
  int ulx, uly, brx, bry;       // UL and BR; UR=(brx,uly), BL=(ulx,bry)
  std::vector<int> keys;        // indices into the candidate list, in vKeys order
  bool no_more = false;
  std::list<OctNode>::iterator self;
};

struct SortEntry { int count; OctNode* node; };

struct Extractor {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // reference member is double (include/ORBextractor.h:92)
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int> quota;
  std::vector<int> umax;
  std::vector<Image> pyr;
  std::vector<std::vector<orc_keypoint>> cand;   // per level, vToDistributeKeys (cell-shifted coords)
  std::vector<std::vector<orc_keypoint>> lvl_kp; // per level after octree+orientation (level coords)
  Image blurred;                 // reused across calls
  std::vector<uint16_t> blur_tmp;

  // ORBextractor.cc:409-469
  Extractor(int nf, float sf, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
    scale.resize(nl); sigma2.resize(nl); inv_scale.resize(nl); inv_sigma2.resize(nl);
    scale[0] = 1.0f; sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
      scale[i] = (float)(scale[i - 1] * scaleFactor);
      sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < nl; i++) {
      inv_scale[i] = 1.0f / scale[i];
      inv_sigma2[i] = 1.0f / sigma2[i];
    }
    pyr.resize(nl);
    quota.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      quota[l] = cv_round(nDesired);
      sum += quota[l];
      nDesired *= factor;
    }
    quota[nl - 1] = std::max(nfeatures - sum, 0);

    umax.assign(HALF_PATCH_SIZE + 2, 0);
    int vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }

  // ORBextractor.cc:1170-1195.  The 19-px REFLECT_101 pad is never read by any
  // stage restated here, so levels are stored unpadded.
  void compute_pyramid(const uint8_t* img, int rows, int cols, int step) {
    for (int l = 0; l < nlevels; l++) {
      float s = inv_scale[l];
      int w = cv_round((float)cols * s), h = cv_round((float)rows * s);
      pyr[l].w = w; pyr[l].h = h;
      pyr[l].px.resize((size_t)w * h);
      if (l == 0) {
        for (int y = 0; y < rows; y++) memcpy(pyr[0].row(y), img + (size_t)y * step, cols);
      } else {
        resize_linear_u8(pyr[l - 1].px.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w,
                         pyr[l].px.data(), w, h, w);
      }
    }
  }

  // ORBextractor.cc:480-536: split into 4 children, points keep parent order.
  static void divide(const OctNode& n, const std::vector<orc_keypoint>& pts, OctNode c[4]) {
    const int halfX = (int)std::ceil(static_cast<float>(n.brx - n.ulx) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(n.bry - n.uly) / 2);
    const int mx = n.ulx + halfX, my = n.uly + halfY;
    c[0].ulx = n.ulx; c[0].uly = n.uly; c[0].brx = mx;    c[0].bry = my;
    c[1].ulx = mx;    c[1].uly = n.uly; c[1].brx = n.brx; c[1].bry = my;
    c[2].ulx = n.ulx; c[2].uly = my;    c[2].brx = mx;    c[2].bry = n.bry;
    c[3].ulx = mx;    c[3].uly = my;    c[3].brx = n.brx; c[3].bry = n.bry;
    for (int i = 0; i < 4; i++) { c[i].keys.clear(); c[i].no_more = false; }
    for (int idx : n.keys) {
      const orc_keypoint& kp = pts[idx];
      if (kp.x < mx) {
        if (kp.y < my) c[0].keys.push_back(idx);
        else c[2].keys.push_back(idx);
      } else if (kp.y < my) c[1].keys.push_back(idx);
      else c[3].keys.push_back(idx);
    }
    for (int i = 0; i < 4; i++)
      if (c[i].keys.size() == 1) c[i].no_more = true;
  }

  // ORBextractor.cc:555-779
  std::vector<orc_keypoint> distribute_octtree(const std::vector<orc_keypoint>& pts, int minX,
                                               int maxX, int minY, int maxY, int N) {
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    // a level more than twice as tall as wide gives nIni == 0: the reference then divides by zero and indexes an empty
    // vector (undefined behaviour); the C ABI rejects such images with ORB_E_ARG and the oracle reports them the same way
    if (nIni < 1) throw std::domain_error("aspect ratio < 0.5: nIni == 0 is undefined in the reference");
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<OctNode> nodes;
    std::vector<OctNode*> ini(nIni);
    for (int i = 0; i < nIni; i++) {
      OctNode n;
      n.ulx = (int)(hX * static_cast<float>(i));
      n.brx = (int)(hX * static_cast<float>(i + 1));
      n.uly = 0;
      n.bry = maxY - minY;
      nodes.push_back(n);
      ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < pts.size(); i++) ini[(int)(pts[i].x / hX)]->keys.push_back((int)i);

    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->keys.size() == 1) { it->no_more = true; ++it; }
      else if (it->keys.empty()) it = nodes.erase(it);
      else ++it;
    }

    auto cmp = [](const SortEntry& a, const SortEntry& b) {  // compareNodes :538-553
      if (a.count < b.count) return true;
      if (a.count > b.count) return false;
      return a.node->ulx < b.node->ulx;
    };
    // push the non-empty children of `parent` to the list front in n1..n4 order
    auto push_children = [&](OctNode c[4], std::vector<SortEntry>& expandable, int* nToExpand) {
      for (int i = 0; i < 4; i++) {
        if (c[i].keys.empty()) continue;
        nodes.push_front(c[i]);
        nodes.front().self = nodes.begin();
        if (c[i].keys.size() > 1) {
          if (nToExpand) (*nToExpand)++;
          expandable.push_back(SortEntry{(int)c[i].keys.size(), &nodes.front()});
        }
      }
    };

    bool finish = false;
    std::vector<SortEntry> expandable;
    while (!finish) {
      int prevSize = (int)nodes.size();
      int nToExpand = 0;
      expandable.clear();
      for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->no_more) { ++it; continue; }
        OctNode c[4];
        divide(*it, pts, c);
        push_children(c, expandable, &nToExpand);
        it = nodes.erase(it);
      }
      if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
        finish = true;
      } else if ((int)nodes.size() + nToExpand * 3 > N) {
        while (!finish) {
          prevSize = (int)nodes.size();
          std::vector<SortEntry> prev = expandable;
          expandable.clear();
          std::sort(prev.begin(), prev.end(), cmp);  // same libstdc++ introsort as the reference
          for (int j = (int)prev.size() - 1; j >= 0; j--) {
            OctNode c[4];
            divide(*prev[j].node, pts, c);
            push_children(c, expandable, nullptr);
            nodes.erase(prev[j].node->self);
            if ((int)nodes.size() >= N) break;
          }
          if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
        }
      }
    }

    std::vector<orc_keypoint> result;
    result.reserve(nodes.size());
    for (auto& n : nodes) {
      int best = n.keys[0];
      float maxResponse = pts[best].response;
      for (size_t k = 1; k < n.keys.size(); k++)
        if (pts[n.keys[k]].response > maxResponse) {
          best = n.keys[k];
          maxResponse = pts[best].response;
        }
      result.push_back(pts[best]);
    }
    return result;
  }

  // ORBextractor.cc:76-103
  float ic_angle(const Image& im, float px, float py) const {
    int m_01 = 0, m_10 = 0;
    const int step = im.w;
    const uint8_t* center = im.row(cv_round(py)) + cv_round(px);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      int v_sum = 0;
      int d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int val_plus = center[u + v * step], val_minus = center[u - v * step];
        v_sum += (val_plus - val_minus);
        m_10 += u * (val_plus + val_minus);
      }
      m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
  }

  // ORBextractor.cc:781-896
  void compute_keypoints() {
    cand.assign(nlevels, {});
    lvl_kp.assign(nlevels, {});
    const float W = 35;
    std::vector<RawKp> cell;
    for (int level = 0; level < nlevels; ++level) {
      const Image& im = pyr[level];
      const int minBorderX = EDGE_THRESHOLD - 3;
      const int minBorderY = minBorderX;
      const int maxBorderX = im.w - EDGE_THRESHOLD + 3;
      const int maxBorderY = im.h - EDGE_THRESHOLD + 3;
      std::vector<orc_keypoint>& toDistribute = cand[level];
      toDistribute.reserve(nfeatures * 10);
      const float width = (float)(maxBorderX - minBorderX);
      const float height = (float)(maxBorderY - minBorderY);
      const int nCols = (int)(width / W);
      const int nRows = (int)(height / W);
      const int wCell = (int)std::ceil(width / nCols);
      const int hCell = (int)std::ceil(height / nRows);
      for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
          const float iniX = (float)(minBorderX + j * wCell);
          float maxX = iniX + wCell + 6;
          if (iniX >= maxBorderX - 6) continue;
          if (maxX > maxBorderX) maxX = (float)maxBorderX;
          const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
          const uint8_t* p = im.row(y0) + x0;
          fast_detect(p, x1 - x0, y1 - y0, im.w, iniTh, true, cell);
          if (cell.empty()) fast_detect(p, x1 - x0, y1 - y0, im.w, minTh, true, cell);
          for (const RawKp& r : cell) {
            orc_keypoint kp;
            kp.x = (float)r.x; kp.y = (float)r.y;
            kp.x += j * wCell;
            kp.y += i * hCell;
            kp.size = 7.f; kp.angle = -1.f; kp.response = (float)r.score;
            kp.octave = 0; kp.class_id = -1;
            toDistribute.push_back(kp);
          }
        }
      }
      std::vector<orc_keypoint>& kps = lvl_kp[level];
      kps = distribute_octtree(toDistribute, minBorderX, maxBorderX, minBorderY, maxBorderY,
                               quota[level]);
      const int scaledPatchSize = (int)(PATCH_SIZE * scale[level]);
      for (auto& kp : kps) {
        kp.x += minBorderX;
        kp.y += minBorderY;
        kp.octave = level;
        kp.size = (float)scaledPatchSize;
      }
    }
    for (int level = 0; level < nlevels; ++level)
      for (auto& kp : lvl_kp[level]) kp.angle = ic_angle(pyr[level], kp.x, kp.y);
  }

  // ORBextractor.cc:107-146
  static void orb_descriptor(const orc_keypoint& kpt, const Image& blurred, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);
    float angle = (float)kpt.angle * factorPI;
    float a = (float)cosf(angle), b = (float)sinf(angle);
    const int step = blurred.w;
    const uint8_t* center = blurred.row(cv_round(kpt.y)) + cv_round(kpt.x);
    const int* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
      int val = 0;
      for (int k = 0; k < 8; k++) {
        const int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
        int t0 = center[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
        int t1 = center[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // ORBextractor.cc:1086-1168.  Returns monoIndex; *n_out = total keypoints.
  int extract(const uint8_t* img, int rows, int cols, int step, int lap0, int lap1,
              orc_keypoint* out_kp, uint8_t* out_desc, int cap, int* n_out) {
    if (!img || rows <= 0 || cols <= 0) return -1;
    compute_pyramid(img, rows, cols, step);
    compute_keypoints();
    int nkeypoints = 0;
    for (int l = 0; l < nlevels; l++) nkeypoints += (int)lvl_kp[l].size();
    *n_out = nkeypoints;
    if (nkeypoints > cap) return -2;
    int monoIndex = 0, stereoIndex = nkeypoints - 1;
    uint8_t d[32];
    for (int l = 0; l < nlevels; l++) {
      std::vector<orc_keypoint>& kps = lvl_kp[l];
      if (kps.empty()) continue;
      blurred.w = pyr[l].w; blurred.h = pyr[l].h;
      blurred.px.resize(pyr[l].px.size());
      gaussian_blur7(pyr[l].px.data(), pyr[l].w, pyr[l].h, pyr[l].w, blurred.px.data(), blurred.w, blur_tmp);
      const float s = scale[l];
      for (const orc_keypoint& src : kps) {
        orb_descriptor(src, blurred, d);
        orc_keypoint kp = src;
        if (l != 0) { kp.x *= s; kp.y *= s; }
        int pos;
        if (kp.x >= lap0 && kp.x <= lap1) pos = stereoIndex--;
        else pos = monoIndex++;
        out_kp[pos] = kp;
        memcpy(out_desc + (size_t)pos * 32, d, 32);
      }
    }
    return monoIndex;
  }
};

}  // namespace

extern "C" {

void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }

int orc_extract(void* h, const uint8_t* img, int rows, int cols, int step, int lap0, int lap1,
                orc_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
  try {
    return ((Extractor*)h)->extract(img, rows, cols, step, lap0, lap1, kps, desc, cap, n_out);
  } catch (const std::domain_error&) {
    return -3;
  }
}

// --- introspection used by the parity tests to localise a mismatch ---
int orc_level_info(void* h, int level, int* w, int* hh, int* quota, float* scale) {
  Extractor* e = (Extractor*)h;
  if (level < 0 || level >= e->nlevels) return -1;
  *w = e->pyr[level].w; *hh = e->pyr[level].h; *quota = e->quota[level]; *scale = e->scale[level];
  return 0;
}
const uint8_t* orc_level_ptr(void* h, int level) { return ((Extractor*)h)->pyr[level].px.data(); }
int orc_level_candidates(void* h, int level, orc_keypoint* out, int cap) {
  auto& v = ((Extractor*)h)->cand[level];
  int n = (int)std::min<size_t>(v.size(), cap);
  if (out) memcpy(out, v.data(), n * sizeof(orc_keypoint));
  return (int)v.size();
}
int orc_level_keypoints(void* h, int level, orc_keypoint* out, int cap) {
  auto& v = ((Extractor*)h)->lvl_kp[level];
  int n = (int)std::min<size_t>(v.size(), cap);
  if (out) memcpy(out, v.data(), n * sizeof(orc_keypoint));
  return (int)v.size();
}
void orc_umax(void* h, int* out16) {
  for (int i = 0; i < 16; i++) out16[i] = ((Extractor*)h)->umax[i];
}

// --- primitives, for pinning against cv2 ---
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh,
                          int dstep) {
  resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
int orc_fast(const uint8_t* img, int w, int h, int step, int threshold, int nonmax, int* xys /*3*cap*/,
             int cap) {
  std::vector<RawKp> v;
  fast_detect(img, w, h, step, threshold, nonmax != 0, v);
  int n = (int)std::min<size_t>(v.size(), cap);
  for (int i = 0; i < n; i++) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
  return (int)v.size();
}
void orc_gaussian_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
  std::vector<uint16_t> tmp;
  gaussian_blur7(src, w, h, sstep, dst, dstep, tmp);
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_cos_sin_deg(float angle_deg, float* c, float* s) {
  const float factorPI = (float)(M_PI / 180.f);
  float a = angle_deg * factorPI;
  *c = cosf(a); *s = sinf(a);
}
void orc_pattern(int* out1024) { memcpy(out1024, kPattern, sizeof(kPattern)); }

// DistributeOctTree alone (ORBextractor.cc:555-779) on an arbitrary candidate
// list: xys = n x {x, y, response}, coordinates relative to minBorder.
int orc_distribute(const int* xys, int n, int band_w, int band_h, int N, int* out_xys, int cap) {
  Extractor e(1000, 1.2f, 8, 20, 7);
  std::vector<orc_keypoint> pts(n);
  for (int i = 0; i < n; i++) {
    pts[i].x = (float)xys[3 * i]; pts[i].y = (float)xys[3 * i + 1]; pts[i].response = (float)xys[3 * i + 2];
    pts[i].size = 7; pts[i].angle = -1; pts[i].octave = 0; pts[i].class_id = -1;
  }
  std::vector<orc_keypoint> r = e.distribute_octtree(pts, 0, band_w, 0, band_h, N);
  for (int i = 0; i < (int)r.size() && i < cap; i++) {
    out_xys[3 * i] = (int)r[i].x; out_xys[3 * i + 1] = (int)r[i].y; out_xys[3 * i + 2] = (int)r[i].response;
  }
  return (int)r.size();
}

// CPU baseline: `nthreads` std::threads, one extractor instance each (the
// reference runs one thread per extractor, Frame.cc:122-125), each extracting
// `iters` frames round-robin from `frames` (nframes x rows x cols, dense).
// Returns elapsed seconds; *total_kp accumulates keypoints so nothing is elided.
double orc_extract_throughput(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                              const uint8_t* frames, int nframes, int rows, int cols, int nthreads,
                              int iters, long long* total_kp) {
  std::vector<Extractor*> ex(nthreads);
  for (auto& e : ex) e = new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  std::vector<long long> kp(nthreads, 0);
  const int cap = nfeatures * 2 + 64 * nlevels;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t]() {
      std::vector<orc_keypoint> k(cap);
      std::vector<uint8_t> d((size_t)cap * 32);
      for (int i = 0; i < iters; i++) {
        int n = 0;
        const uint8_t* img = frames + (size_t)((t + i) % nframes) * rows * cols;
        ex[t]->extract(img, rows, cols, cols, 0, 0, k.data(), d.data(), cap, &n);
        kp[t] += n;
      }
    });
  for (auto& x : th) x.join();
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  long long tot = 0;
  for (int t = 0; t < nthreads; t++) { tot += kp[t]; delete ex[t]; }
  if (total_kp) *total_kp = tot;
  return dt;
}

// std::sort with the reference's comparator on (count, ulx) pairs; returns the
// permutation -- used to pin the product's re-implementation of libstdc++'s
// introsort tie behaviour (ORBextractor.cc:700).
void orc_sort_nodes(const int* count, const int* ulx, int n, int* perm_out) {
  struct E { int count, ulx, id; };
  std::vector<E> v(n);
  for (int i = 0; i < n; i++) v[i] = E{count[i], ulx[i], i};
  std::sort(v.begin(), v.end(), [](const E& a, const E& b) {
    if (a.count < b.count) return true;
    if (a.count > b.count) return false;
    return a.ulx < b.ulx;
  });
  for (int i = 0; i < n; i++) perm_out[i] = v[i].id;
}

}  // extern "C"
