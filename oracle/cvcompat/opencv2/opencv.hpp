// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the OpenCV C++ API, just wide enough to compile the
// reference's src/ORBextractor.cc UNMODIFIED into oracle/_ref/ (OpenCV's C++ headers and libraries are not
// in this image).  Nothing in the product includes this file.
//
// What it is: containers and value types (Mat with ROI semantics, KeyPoint, Point_, Size, Rect, the
// InputArray / OutputArray proxies) written from OpenCV's documented behaviour, and the image primitives the
// extractor calls.  The un-vendored arithmetic (resize INTER_LINEAR, FAST-9/16 + NMS, GaussianBlur 7x7 s=2,
// fastAtan2) is NOT re-derived here: those entry points forward to the oracle's primitives
// (oracle/orc_extract.cpp: orc_resize_linear_u8, orc_fast, orc_gaussian_blur7, orc_fast_atan2), which
// tests/test_oracle_vs_cv2.py pins bit-exact to cv2 4.13.  So oracle/_ref = the reference's own control flow
// (pyramid, cell loop + fallback, DivideNode / DistributeOctTree / std::list + std::sort, IC_Angle,
// computeOrbDescriptor, operator()'s ordering) as compiled object code + cv2-pinned leaves.
// Anything the extractor does not use is absent on purpose; unsupported uses abort loudly.
#pragma once
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <climits>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

extern "C" {
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep);
int orc_fast(const uint8_t* img, int w, int h, int step, int threshold, int nonmax, int* xys, int cap);
void orc_gaussian_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep);
float orc_fast_atan2(float y, float x);
}

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5  /* named by DBoW2's FORB::toMat32F (never called here); Mat elements are bytes */
#define CVCOMPAT_DIE(msg)                                                     \
  do {                                                                        \
    fprintf(stderr, "cvcompat: unsupported use: %s (%s:%d)\n", msg, __FILE__, __LINE__); \
    abort();                                                                  \
  } while (0)

typedef unsigned char uchar;

// cvRound: round half to even (SSE cvtss2si / lrint), cvFloor / cvCeil as documented
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

typedef ::uchar uchar;

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}  // arguments convert like any T parameter (float -> int truncates)
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
template <>
inline Point_<float>& Point_<float>::operator*=(float s) { x *= s; y *= s; return *this; }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

class KeyPoint {
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

struct MatStep {
  size_t v;
  MatStep() : v(0) {}
  operator size_t() const { return v; }
};

// 8-bit single-channel matrix header over shared storage; sub-matrix headers remember their place in the
// allocation (locateROI) because copyMakeBorder without BORDER_ISOLATED reads the parent's pixels.
class Mat {
 public:
  int rows, cols;
  uchar* data;
  MatStep step;
  Mat() : rows(0), cols(0), data(nullptr), whole_rows(0), whole_cols(0), ofs_x(0), ofs_y(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* user, size_t user_step) : Mat() {  // header over caller memory
    if (type != CV_8UC1) CVCOMPAT_DIE("only CV_8UC1");
    rows = r; cols = c; data = (uchar*)user; step.v = user_step;
    whole_rows = r; whole_cols = c; ofs_x = ofs_y = 0;
  }
  void create(int r, int c, int type) {
    if (type != CV_8UC1) CVCOMPAT_DIE("only CV_8UC1");
    if (data && r == rows && c == cols) return;  // cv::Mat::create keeps a matching allocation
    buf = std::shared_ptr<uchar>(new uchar[(size_t)std::max(r, 0) * std::max(c, 0) + 1], std::default_delete<uchar[]>());
    rows = r; cols = c; data = buf.get(); step.v = (size_t)c;
    whole_rows = r; whole_cols = c; ofs_x = ofs_y = 0;
  }
  void release() { buf.reset(); rows = cols = 0; data = nullptr; step.v = 0; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  size_t step1() const { return step.v; }
  Size size() const { return Size(cols, rows); }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
  Mat operator()(const Rect& r) const {
    if (r.x < 0 || r.y < 0 || r.x + r.width > cols || r.y + r.height > rows) CVCOMPAT_DIE("ROI outside the matrix");
    Mat m(*this);
    m.rows = r.height; m.cols = r.width; m.data = data + (size_t)r.y * step.v + r.x;
    m.ofs_x = ofs_x + r.x; m.ofs_y = ofs_y + r.y;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), cols);
    return m;
  }
  inline void copyTo(OutputArray dst) const;
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    if (m.data) memset(m.data, 0, (size_t)r * c);
    return m;
  }
  void locateROI(Size& whole, Point& ofs) const { whole = Size(whole_cols, whole_rows); ofs = Point(ofs_x, ofs_y); }
  bool isSubmatrix() const { return rows != whole_rows || cols != whole_cols; }
  // ---- named by src/Frame.cc / ORBmatcher.cc / MapPoint.cc (oracle/_ref/libref_front.so) on paths that build float
  //      matrices (calibration, serialisation, drawing): they type-check, and abort if a run ever reaches them
  Mat(int r, int c, int type, void* user) : Mat() { (void)r; (void)c; (void)type; (void)user; CVCOMPAT_DIE("Mat over user data of another type"); }
  template <typename T> T& at(int i) { return *(T*)(data + (size_t)i * sizeof(T)); }
  template <typename T> const T& at(int i) const { return *(const T*)(data + (size_t)i * sizeof(T)); }
  size_t elemSize() const { return 1; }
  size_t total() const { return (size_t)rows * cols; }
  bool isContinuous() const { return step.v == (size_t)cols; }
  Mat reshape(int, int = 0) const { CVCOMPAT_DIE("reshape"); return Mat(); }
  Mat t() const { CVCOMPAT_DIE("t"); return Mat(); }
  Mat inv(int = 0) const { CVCOMPAT_DIE("inv"); return Mat(); }
  Mat col(int) const { CVCOMPAT_DIE("col"); return Mat(); }
  void convertTo(const Mat&, int, double = 1, double = 0) const { CVCOMPAT_DIE("convertTo"); }
  int channels() const { return 1; }
  int depth() const { return 0; }
  static Mat eye(int r, int c, int type) { (void)r; (void)c; (void)type; CVCOMPAT_DIE("eye"); return Mat(); }
  static Mat ones(int r, int c, int type) { (void)r; (void)c; (void)type; CVCOMPAT_DIE("ones"); return Mat(); }

 private:
  std::shared_ptr<uchar> buf;
  int whole_rows, whole_cols, ofs_x, ofs_y;
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  Mat getMat() const { return *m_; }
  bool empty() const { return m_->empty(); }

 protected:
  const Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray(Mat& m) : _InputArray(m) {}
  _OutputArray(const Mat& m) : _InputArray(m) {}  // a temporary header (e.g. descriptors.row(i)): written through
  void create(int r, int c, int type) const { const_cast<Mat*>(m_)->create(r, c, type); }
  void create(Size sz, int type) const { create(sz.height, sz.width, type); }
  void release() const { const_cast<Mat*>(m_)->release(); }
};

inline void Mat::copyTo(OutputArray dst) const {
  dst.create(rows, cols, CV_8UC1);
  Mat d = dst.getMat();
  for (int y = 0; y < rows; y++) memmove(d.ptr(y), ptr(y), cols);
}

// ---- primitives whose arithmetic lives in the cv2-pinned oracle ------------------------------------------
inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
  const Mat img = image.getMat();
  keypoints.clear();
  if (img.empty()) return;
  std::vector<int> xys((size_t)3 * img.rows * img.cols + 3);
  const int n = orc_fast(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0, xys.data(),
                         img.rows * img.cols);
  keypoints.reserve(n);
  // cv::FAST emits KeyPoint(x, y, 7.f, -1, score) row-major
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}

inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
  if (fx != 0 || fy != 0 || interpolation != INTER_LINEAR) CVCOMPAT_DIE("resize: only dsize + INTER_LINEAR");
  const Mat src = _src.getMat();
  _dst.create(dsize, src.type());
  Mat dst = _dst.getMat();
  orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
  if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 || (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101)
    CVCOMPAT_DIE("GaussianBlur: only 7x7, sigma 2, BORDER_REFLECT_101");
  const Mat src = _src.getMat();
  if (src.isSubmatrix() && !(borderType & BORDER_ISOLATED)) CVCOMPAT_DIE("GaussianBlur on a sub-matrix reads its parent");
  const Mat in = src.clone();  // in-place calls are allowed
  _dst.create(src.size(), src.type());
  Mat dst = _dst.getMat();
  orc_gaussian_blur7(in.data, in.cols, in.rows, (int)in.step, dst.data, (int)dst.step);
}

// copyMakeBorder, BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba).  Without BORDER_ISOLATED a sub-matrix source is
// first widened into its parent by up to the border widths (OpenCV reads the real neighbours), the rest is
// extrapolated.  The in-place form of ComputePyramid (src = centre ROI of dst) only writes border pixels.
inline int borderInterpolate101(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}
inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) CVCOMPAT_DIE("copyMakeBorder: only BORDER_REFLECT_101");
  Mat src = _src.getMat();
  const int out_rows = src.rows + top + bottom, out_cols = src.cols + left + right;
  const uchar* sdata = src.data;
  int srows = src.rows, scols = src.cols;
  if (src.isSubmatrix() && !(borderType & BORDER_ISOLATED)) {
    Size whole; Point ofs;
    src.locateROI(whole, ofs);
    const int dtop = std::min(ofs.y, top), dbottom = std::min(whole.height - src.rows - ofs.y, bottom);
    const int dleft = std::min(ofs.x, left), dright = std::min(whole.width - src.cols - ofs.x, right);
    sdata -= (size_t)dtop * src.step + dleft;
    srows += dtop + dbottom; scols += dleft + dright;
    top -= dtop; left -= dleft;
  }
  _dst.create(out_rows, out_cols, src.type());
  Mat dst = _dst.getMat();
  const size_t sstep = src.step;
  for (int y = 0; y < out_rows; y++) {
    const uchar* srow = sdata + (size_t)borderInterpolate101(y - top, srows) * sstep;
    uchar* drow = dst.ptr(y);
    const bool interior_row = (y - top) >= 0 && (y - top) < srows;
    for (int x = 0; x < out_cols; x++) {
      const int sx = x - left;
      if (interior_row && sx >= 0 && sx < scols) {
        if (drow + x != srow + sx) drow[x] = srow[sx];
      } else {
        drow[x] = srow[borderInterpolate101(sx, scols)];
      }
    }
  }
}

// only referenced by the (uncalled) ComputeKeyPointsOld: keep the n strongest responses, ties at the cut kept
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
    if (npoints >= 0 && keypoints.size() > (size_t)npoints) {
      if (npoints == 0) { keypoints.clear(); return; }
      std::nth_element(keypoints.begin(), keypoints.begin() + npoints - 1, keypoints.end(),
                       [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
      const float amb = keypoints[npoints - 1].response;
      auto end = std::partition(keypoints.begin() + npoints, keypoints.end(), [amb](const KeyPoint& k) { return k.response >= amb; });
      keypoints.resize(end - keypoints.begin());
    }
  }
};

// ---- types and functions the front-end sources name outside the paths oracle/_ref runs (type-check only)
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f; typedef Point3_<double> Point3d; typedef Point_<double> Point2d;
template <typename T> struct Mat_ : Mat {
  Mat_() {} Mat_(int, int) { CVCOMPAT_DIE("Mat_"); } Mat_(const Mat&) {}
  T& operator()(int, int) { static T t; CVCOMPAT_DIE("Mat_()"); return t; } T& operator()(int) { static T t; CVCOMPAT_DIE("Mat_()"); return t; }
};
template <typename T> struct MatCommaInitializer_ { template <class U> MatCommaInitializer_& operator,(const U&) { return *this; } operator Mat() const { return Mat(); } operator Mat_<T>() const { return Mat_<T>(); } };
template <typename T, class U> MatCommaInitializer_<T> operator<<(const Mat_<T>&, const U&) { CVCOMPAT_DIE("Mat_ <<"); return MatCommaInitializer_<T>(); }
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {} };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
struct BFMatcher { BFMatcher(int = 4, bool = false) {} template <class... A> void match(const A&...) const { CVCOMPAT_DIE("BFMatcher"); } template <class... A> void knnMatch(const A&...) const { CVCOMPAT_DIE("BFMatcher"); } };
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, CV_32FC1 = 5, CV_64F = 6 };
inline std::ostream& operator<<(std::ostream& o, const Mat&) { return o; }
inline Mat operator*(const Mat&, const Mat&) { CVCOMPAT_DIE("Mat * Mat"); return Mat(); }
inline Mat operator+(const Mat&, const Mat&) { CVCOMPAT_DIE("Mat + Mat"); return Mat(); }
inline Mat operator-(const Mat&, const Mat&) { CVCOMPAT_DIE("Mat - Mat"); return Mat(); }
inline Mat operator-(const Mat&) { CVCOMPAT_DIE("-Mat"); return Mat(); }
inline Mat operator*(const Mat&, double) { CVCOMPAT_DIE("Mat * s"); return Mat(); }
inline Mat operator/(const Mat&, double) { CVCOMPAT_DIE("Mat / s"); return Mat(); }
template <class... A> void undistortPoints(const A&...) { CVCOMPAT_DIE("undistortPoints"); }
template <class... A> void hconcat(const A&...) { CVCOMPAT_DIE("hconcat"); }
template <class... A> void vconcat(const A&...) { CVCOMPAT_DIE("vconcat"); }
namespace fisheye { template <class... A> void undistortPoints(const A&...) { CVCOMPAT_DIE("fisheye::undistortPoints"); } }
// cv::norm(a, b, NORM_L1) of two 8-bit matrices of the same size: the SAD of Frame::ComputeStereoMatches (Frame.cc:908-923),
// an exact integer sum returned as double like cv::norm does
inline double norm(const Mat& a, const Mat& b, int type) {
  if (type != NORM_L1 || a.rows != b.rows || a.cols != b.cols) CVCOMPAT_DIE("norm: only L1 of equal sizes");
  long long s = 0;
  for (int y = 0; y < a.rows; y++) { const uchar* p = a.ptr(y); const uchar* q = b.ptr(y); for (int x = 0; x < a.cols; x++) s += p[x] > q[x] ? p[x] - q[x] : q[x] - p[x]; }
  return (double)s;
}
template <typename T> double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
template <typename T> double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }

// cv::FileStorage / FileNode: named by DBoW2's TemplatedVocabulary::save / load (YAML), which oracle/_ref never calls
// (the vocabulary comes in through loadFromTextFile): declarations that type-check, nothing more.
struct FileNodeIterator;
struct FileNode {
  FileNode operator[](const std::string&) const { return FileNode(); } FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  bool empty() const { return true; } int type() const { return 0; } size_t size() const { return 0; }
  operator int() const { return 0; } operator float() const { return 0; } operator double() const { return 0; } operator std::string() const { return std::string(); }
  enum { NONE = 0, INT = 1, REAL = 2, STRING = 3, SEQ = 4, MAP = 5 };
};
template <class T> void operator>>(const FileNode&, T&) {}
struct FileStorage {
  enum { READ = 0, WRITE = 1 };
  FileStorage() {} FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; } void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); } FileNode operator[](const char*) const { return FileNode(); }
  FileNode getFirstTopLevelNode() const { return FileNode(); }
};
template <class T> FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
