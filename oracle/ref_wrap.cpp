// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's own object code (oracle/_ref build):
// /root/reference/src/ORBextractor.cc compiled unmodified against oracle/cvcompat/, and the two free-standing
// members of /root/reference/src/ORBmatcher.cc (ComputeThreeMaxima :2012-2054, DescriptorDistance :2059-2074)
// that need nothing but cv::Mat.  The rest of ORBmatcher.cc and all of g2o need Eigen / Sophus / DBoW2 headers
// that are not in this image: those rows stay pinned by the numpy restatements.
// Nothing in the product links or loads this library.
#include <chrono>
#include <thread>
#include <vector>

#include "ORBextractor.h"  // the reference's header (-I /root/reference/include)
#include "orc_common.h"

namespace ORB_SLAM3 {
// the two members this file exercises, declared as in include/ORBmatcher.h:43 and :92 (the full header needs Sophus)
class ORBmatcher {
 public:
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
  static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
};
}  // namespace ORB_SLAM3

namespace {
struct Probe : public ORB_SLAM3::ORBextractor {  // read-only view of the protected tables
  using ORB_SLAM3::ORBextractor::ORBextractor;
  const std::vector<int>& quotas() const { return mnFeaturesPerLevel; }
  const std::vector<int>& um() const { return umax; }
  const std::vector<cv::Point>& pat() const { return pattern; }
};
static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "KeyPoint layout (SURVEY.md A.7)");
}  // namespace

extern "C" {

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new Probe(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_extractor_destroy(void* h) { delete (Probe*)h; }

// ORBextractor::operator() (ORBextractor.cc:1086-1168).  Returns monoIndex (or -1), *n_out = keypoints.
int ref_extract(void* h, const uint8_t* img, int rows, int cols, int step, int lap0, int lap1, orc_keypoint* kps,
                uint8_t* desc, int cap, int* n_out) {
  Probe* e = (Probe*)h;
  cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step), descriptors;
  if (!img) image = cv::Mat();
  std::vector<cv::KeyPoint> keys;
  std::vector<int> lap = {lap0, lap1};
  const int mono = (*e)(image, cv::Mat(), keys, descriptors, lap);
  const int n = (int)keys.size();
  if (n_out) *n_out = n;
  for (int i = 0; i < n && i < cap; i++) {
    memcpy(&kps[i], &keys[i], sizeof(orc_keypoint));
    memcpy(desc + (size_t)32 * i, descriptors.ptr(i), 32);
  }
  return mono;
}

int ref_level_info(void* h, int level, int* w, int* hh, int* step, int* quota, float* scale) {
  Probe* e = (Probe*)h;
  if (level < 0 || level >= e->GetLevels() || e->mvImagePyramid[level].empty()) return -1;
  const cv::Mat& m = e->mvImagePyramid[level];
  *w = m.cols; *hh = m.rows; *step = (int)m.step; *quota = e->quotas()[level]; *scale = e->GetScaleFactors()[level];
  return 0;
}
// (x, y) may reach 19 px outside the level: the EDGE_THRESHOLD border of ComputePyramid (:1185-1191)
const uint8_t* ref_level_ptr(void* h, int level) { return ((Probe*)h)->mvImagePyramid[level].data; }
void ref_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota, int* umax16,
                int* pattern1024) {
  Probe* e = (Probe*)h;
  const int nl = e->GetLevels();
  for (int i = 0; i < nl; i++) {
    scale[i] = e->GetScaleFactors()[i]; inv_scale[i] = e->GetInverseScaleFactors()[i];
    sigma2[i] = e->GetScaleSigmaSquares()[i]; inv_sigma2[i] = e->GetInverseScaleSigmaSquares()[i];
    quota[i] = e->quotas()[i];
  }
  for (int i = 0; i < 16; i++) umax16[i] = e->um()[i];
  for (int i = 0; i < 512; i++) { pattern1024[2 * i] = e->pat()[i].x; pattern1024[2 * i + 1] = e->pat()[i].y; }
}

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  alignas(4) uint8_t ta[32], tb[32];
  memcpy(ta, a, 32); memcpy(tb, b, 32);
  const cv::Mat ma(1, 32, CV_8UC1, ta, 32), mb(1, 32, CV_8UC1, tb, 32);
  return ORB_SLAM3::ORBmatcher::DescriptorDistance(ma, mb);
}
// histo sizes -> the three maxima (the reference only looks at histo[i].size())
void ref_three_maxima(const int* sizes, int L, int* ind /*3, in: initial values*/) {
  std::vector<std::vector<int>> h(L);
  for (int i = 0; i < L; i++) h[i].resize(sizes[i]);
  ORB_SLAM3::ORBmatcher m;
  m.ComputeThreeMaxima(h.data(), L, ind[0], ind[1], ind[2]);
}

// CPU baseline of bench.py (kind "reference"): `nthreads` extractor instances, one std::thread each, exactly
// like orc_extract_throughput of the restatement.
double ref_extract_throughput(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* frames,
                              int nframes, int rows, int cols, int nthreads, int iters, long long* total_kp) {
  std::vector<Probe*> ex(nthreads);
  for (auto& e : ex) e = new Probe(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  std::vector<long long> kp(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t]() {
      std::vector<cv::KeyPoint> keys;
      std::vector<int> lap = {0, 0};
      for (int i = 0; i < iters; i++) {
        cv::Mat image(rows, cols, CV_8UC1, (void*)(frames + (size_t)((t + i) % nframes) * rows * cols), (size_t)cols), d;
        (*ex[t])(image, cv::Mat(), keys, d, lap);
        kp[t] += (long long)keys.size();
      }
    });
  for (auto& x : th) x.join();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  long long tot = 0;
  for (int t = 0; t < nthreads; t++) { tot += kp[t]; delete ex[t]; }
  if (total_kp) *total_kp = tot;
  return dt;
}

}  // extern "C"
