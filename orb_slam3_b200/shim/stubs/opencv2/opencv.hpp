// SYNTAX-CHECK STAND-IN, not OpenCV: the declarations the reference's headers and this repo's shims use, so that
// `g++ -fsyntax-only` can hold the shims against the reference's own, unmodified headers in an image without the
// OpenCV C++ headers (tests/test_shim_syntax.py).  Nothing here computes anything useful; a real build uses the real
// OpenCV.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <cstddef>
#include <cstdint>
#include <list>
#include <memory>
#include <ostream>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_32S 4
#define CV_PI 3.1415926535897932384626433832795
typedef unsigned char uchar;
namespace cv {
template <class T> struct Point_ {
  T x = T(), y = T();
  Point_() {}
  Point_(T a, T b) : x(a), y(b) {}
  template <class U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
  Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
  template <class U> Point_ operator*(U s) const { return Point_((T)(x * s), (T)(y * s)); }
  bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
typedef Point_<int> Point2i; typedef Point2i Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <class T> struct Point3_ {
  T x = T(), y = T(), z = T();
  Point3_() {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
  Point3_ operator-(const Point3_& o) const { return Point3_(x - o.x, y - o.y, z - o.z); }
  Point3_ operator+(const Point3_& o) const { return Point3_(x + o.x, y + o.y, z + o.z); }
};
typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;
template <class T> struct Size_ { T width = T(), height = T(); Size_() {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
template <class T> struct Rect_ { T x = T(), y = T(), width = T(), height = T(); Rect_() {} Rect_(T a, T b, T c, T d) : x(a), y(b), width(c), height(d) {} };
typedef Rect_<int> Rect;
struct Range { int start = 0, end = 0; Range() {} Range(int a, int b) : start(a), end(b) {} static Range all() { return Range(); } };
template <class T, int N> struct Vec { T val[N]; T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } T& operator()(int i) { return val[i]; } };
typedef Vec<float, 3> Vec3f; typedef Vec<double, 3> Vec3d; typedef Vec<uchar, 3> Vec3b; typedef Vec<float, 4> Vec4f; typedef Vec<float, 2> Vec2f;
template <class T, int M, int N> struct Matx { T val[M * N]; T& operator()(int i, int j) { return val[i * N + j]; } const T& operator()(int i, int j) const { return val[i * N + j]; } };
typedef Matx<float, 3, 3> Matx33f; typedef Matx<double, 3, 3> Matx33d; typedef Matx<float, 3, 1> Matx31f; typedef Matx<float, 4, 4> Matx44f;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {} };
struct KeyPoint {
  Point2f pt; float size = 0; float angle = -1; float response = 0; int octave = 0; int class_id = -1;
  KeyPoint() {}
  KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
struct BFMatcher { BFMatcher(int = 4, bool = false) {} template <class... A> void match(const A&...) const {} template <class... A> void knnMatch(const A&...) const {} };
struct MatExpr;
struct Mat {
  uint8_t* data = nullptr; int rows = 0, cols = 0, dims = 2; struct Step { size_t v = 0; operator size_t() const { return v; } size_t operator[](int) const { return v; } Step& operator=(size_t s) { v = s; return *this; } } step;
  Mat() {}
  Mat(int, int, int) {}
  Mat(int, int, int, const Scalar&) {}
  Mat(Size, int) {}
  Mat(int, int, int, void*, size_t = 0) {}
  template <class T> explicit Mat(const std::vector<T>&, bool = false) {}
  template <class T, int M, int N> Mat(const Matx<T, M, N>&) {}
  Mat(const Mat&, const Rect&) {}
  Mat(const Mat&, const Range&, const Range& = Range::all()) {}
  Mat(const MatExpr&) {}
  Mat& operator=(const MatExpr&) { return *this; }
  Mat& operator=(const Scalar&) { return *this; }
  int type() const { return CV_8UC1; } int depth() const { return 0; } int channels() const { return 1; }
  bool empty() const { return !data; } bool isContinuous() const { return true; } size_t total() const { return 0; } size_t elemSize() const { return 1; }
  Size size() const { return Size(cols, rows); }
  Mat rowRange(int, int) const { return Mat(); } Mat colRange(int, int) const { return Mat(); } Mat row(int) const { return Mat(); } Mat col(int) const { return Mat(); }
  Mat rowRange(const Range&) const { return Mat(); } Mat colRange(const Range&) const { return Mat(); }
  Mat clone() const { return Mat(); } Mat t() const { return Mat(); } Mat inv(int = 0) const { return Mat(); } Mat mul(const Mat&, double = 1) const { return Mat(); }
  Mat reshape(int, int = 0) const { return Mat(); } Mat operator()(const Rect&) const { return Mat(); } Mat operator()(const Range&, const Range&) const { return Mat(); }
  Mat diag(int = 0) const { return Mat(); } Mat cross(const Mat&) const { return Mat(); } double dot(const Mat&) const { return 0; }
  void copyTo(const Mat&) const {} void copyTo(const Mat&, const Mat&) const {} void convertTo(const Mat&, int, double = 1, double = 0) const {}
  void release() {} void create(int, int, int) {} void create(Size, int) {} void push_back(const Mat&) {} template <class T> void push_back(const T&) {}
  Mat& setTo(const Scalar&) { return *this; }
  template <class T> T& at(int) { static T t; return t; } template <class T> T& at(int, int) { static T t; return t; }
  template <class T> const T& at(int) const { static T t; return t; } template <class T> const T& at(int, int) const { static T t; return t; }
  template <class T> T& at(Point) { static T t; return t; }
  template <class T> T* ptr(int = 0) { return nullptr; } template <class T> const T* ptr(int = 0) const { return nullptr; }
  uchar* ptr(int = 0) { return nullptr; } const uchar* ptr(int = 0) const { return nullptr; }
  static Mat zeros(int, int, int) { return Mat(); } static Mat ones(int, int, int) { return Mat(); } static Mat eye(int, int, int) { return Mat(); }
  static Mat zeros(Size, int) { return Mat(); }
};
struct MatExpr { MatExpr() {} MatExpr(const Mat&) {} operator Mat() const { return Mat(); } Mat t() const { return Mat(); } Mat inv(int = 0) const { return Mat(); } };
template <class T> struct Mat_ : Mat { Mat_() {} Mat_(int, int) {} template <class U> Mat_(const U&) {} T& operator()(int, int) { static T t; return t; } T& operator()(int) { static T t; return t; } };
template <class T> struct MatCommaInitializer_ { template <class U> MatCommaInitializer_& operator,(const U&) { return *this; } operator Mat() const { return Mat(); } operator Mat_<T>() const { return Mat_<T>(); } };
template <class T, class U> MatCommaInitializer_<T> operator<<(const Mat_<T>&, const U&) { return MatCommaInitializer_<T>(); }
inline MatExpr operator*(const Mat&, const Mat&) { return MatExpr(); } inline MatExpr operator+(const Mat&, const Mat&) { return MatExpr(); }
inline MatExpr operator-(const Mat&, const Mat&) { return MatExpr(); } inline MatExpr operator-(const Mat&) { return MatExpr(); }
inline MatExpr operator*(const Mat&, double) { return MatExpr(); } inline MatExpr operator*(double, const Mat&) { return MatExpr(); }
inline MatExpr operator/(const Mat&, double) { return MatExpr(); } inline MatExpr operator*(const MatExpr&, const Mat&) { return MatExpr(); }
inline MatExpr operator*(const Mat&, const MatExpr&) { return MatExpr(); } inline MatExpr operator+(const MatExpr&, const Mat&) { return MatExpr(); }
inline MatExpr operator-(const MatExpr&, const Mat&) { return MatExpr(); } inline MatExpr operator+(const Mat&, const MatExpr&) { return MatExpr(); }
inline MatExpr operator-(const Mat&, const MatExpr&) { return MatExpr(); } inline MatExpr operator*(const MatExpr&, const MatExpr&) { return MatExpr(); }
inline std::ostream& operator<<(std::ostream& o, const Mat&) { return o; }
template <class T> std::ostream& operator<<(std::ostream& o, const Point_<T>&) { return o; }
template <class T> std::ostream& operator<<(std::ostream& o, const Point3_<T>&) { return o; }
struct _InputArray { _InputArray() {} _InputArray(const Mat&) {} _InputArray(const MatExpr&) {} template <class T> _InputArray(const std::vector<T>&) {}
  bool empty() const { return true; } Mat getMat(int = -1) const { return Mat(); } };
struct _OutputArray : _InputArray { _OutputArray() {} _OutputArray(Mat&) {} template <class T> _OutputArray(std::vector<T>&) {}
  void release() const {} void create(int, int, int) const {} void create(Size, int) const {} };
typedef const _InputArray& InputArray; typedef const _OutputArray& OutputArray; typedef const _OutputArray& InputOutputArray;
typedef InputArray InputArrayOfArrays; typedef OutputArray OutputArrayOfArrays;
inline InputArray noArray() { static _InputArray a; return a; }
template <class T> struct Ptr : std::shared_ptr<T> { Ptr() {} template <class U> Ptr(U* p) : std::shared_ptr<T>(p) {} template <class U> Ptr(const std::shared_ptr<U>& p) : std::shared_ptr<T>(p) {} bool empty() const { return !this->get(); } };
struct FileNodeIterator;
struct FileNode {
  FileNode() {}
  FileNode operator[](const std::string&) const { return FileNode(); } FileNode operator[](const char*) const { return FileNode(); } FileNode operator[](int) const { return FileNode(); }
  bool empty() const { return true; } bool isNone() const { return true; } bool isInt() const { return false; } bool isReal() const { return false; } bool isString() const { return false; }
  bool isSeq() const { return false; } bool isMap() const { return false; } int type() const { return 0; } size_t size() const { return 0; } std::string name() const { return ""; }
  operator int() const { return 0; } operator float() const { return 0; } operator double() const { return 0; } operator std::string() const { return ""; }
  Mat mat() const { return Mat(); } double real() const { return 0; } std::string string() const { return ""; }
  FileNodeIterator begin() const; FileNodeIterator end() const;
  enum { SEQ = 4, MAP = 5, NONE = 0, INT = 1, REAL = 2, STRING = 3 };
};
struct FileNodeIterator { FileNode operator*() const { return FileNode(); } FileNodeIterator& operator++() { return *this; } FileNodeIterator operator++(int) { return *this; }
  bool operator!=(const FileNodeIterator&) const { return false; } bool operator==(const FileNodeIterator&) const { return true; } };
inline FileNodeIterator FileNode::begin() const { return FileNodeIterator(); } inline FileNodeIterator FileNode::end() const { return FileNodeIterator(); }
template <class T> void operator>>(const FileNode&, T&) {}
struct FileStorage {
  enum { READ = 0, WRITE = 1, APPEND = 2, MEMORY = 4 };
  FileStorage() {} FileStorage(const std::string&, int, const std::string& = std::string()) {}
  bool open(const std::string&, int, const std::string& = std::string()) { return false; } bool isOpened() const { return false; } void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); } FileNode operator[](const char*) const { return FileNode(); } FileNode root(int = 0) const { return FileNode(); }
  FileNode getFirstTopLevelNode() const { return FileNode(); }
};
template <class T> FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
inline int cvRound(double v) { return (int)v; } inline int cvFloor(double v) { return (int)v; } inline int cvCeil(double v) { return (int)v; }
inline float fastAtan2(float, float) { return 0; }
inline double norm(InputArray, int = 4) { return 0; } inline double norm(InputArray, InputArray, int = 4) { return 0; }
template <class T> double norm(const Point_<T>&) { return 0; } template <class T> double norm(const Point3_<T>&) { return 0; }
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1, DECOMP_SVD = 1, COLOR_RGB2GRAY = 7, COLOR_BGR2GRAY = 6,
       COLOR_GRAY2BGR = 8, COLOR_RGBA2GRAY = 11, COLOR_BGRA2GRAY = 10, FONT_HERSHEY_PLAIN = 1, SVD_FULL_UV = 4, SVD_MODIFY_A = 1 };
struct SVD { enum { MODIFY_A = 1, FULL_UV = 4 }; static void compute(InputArray, OutputArray, OutputArray, OutputArray, int = 0) {} };
inline void resize(InputArray, OutputArray, Size, double = 0, double = 0, int = 1) {}
inline void copyMakeBorder(InputArray, OutputArray, int, int, int, int, int, const Scalar& = Scalar()) {}
inline void GaussianBlur(InputArray, OutputArray, Size, double, double = 0, int = 4) {}
inline void FAST(InputArray, std::vector<KeyPoint>&, int, bool = true) {}
inline void cvtColor(InputArray, OutputArray, int, int = 0) {}
inline void undistortPoints(InputArray, OutputArray, InputArray, InputArray, InputArray = noArray(), InputArray = noArray()) {}
inline void initUndistortRectifyMap(InputArray, InputArray, InputArray, InputArray, Size, int, OutputArray, OutputArray) {}
inline void remap(InputArray, OutputArray, InputArray, InputArray, int, int = 0, const Scalar& = Scalar()) {}
inline void Rodrigues(InputArray, OutputArray, OutputArray = _OutputArray()) {}
inline void hconcat(InputArray, InputArray, OutputArray) {} inline void vconcat(InputArray, InputArray, OutputArray) {}
inline bool solve(InputArray, InputArray, OutputArray, int = 0) { return true; }
inline double determinant(InputArray) { return 0; }
inline void eigen2cv_dummy() {}
template <class T, class M> void eigen2cv(const M&, T&) {}
template <class T, class M> void cv2eigen(const T&, M&) {}
namespace fisheye { inline void undistortPoints(InputArray, OutputArray, InputArray, InputArray, InputArray = noArray(), InputArray = noArray()) {} }
}  // namespace cv
