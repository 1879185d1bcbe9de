// Minimal stand-in for the OpenCV declarations the reference's
// include/ORBextractor.h and shim/ORBextractor.cc use, ONLY so the shim can be
// syntax-checked in an image without OpenCV headers (tests/test_shim_syntax.py).
// A real build uses the real OpenCV.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <list>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
typedef unsigned char uchar;
namespace cv {
struct Point2i { int x = 0, y = 0; Point2i() {} Point2i(int a, int b) : x(a), y(b) {} };
typedef Point2i Point;
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; float size; float angle; float response; int octave; int class_id; };
struct Mat {
  uint8_t* data = nullptr; int rows = 0, cols = 0; size_t step = 0;
  Mat() {}
  Mat(int r, int c, int) : rows(r), cols(c), step(c) { storage.resize((size_t)r * c); data = storage.data(); }
  Mat(int r, int c, int, void* p, size_t s) : data((uint8_t*)p), rows(r), cols(c), step(s) {}
  int type() const { return CV_8UC1; }
  bool empty() const { return !data; }
  Mat rowRange(int a, int b) const { Mat m; m.data = data + a * step; m.rows = b - a; m.cols = cols; m.step = step; return m; }
  void copyTo(Mat m) const { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.data[r * m.step + c] = data[r * step + c]; }
  std::vector<uint8_t> storage;
};
struct _InputArray { const Mat* m = nullptr; _InputArray() {} _InputArray(const Mat& x) : m(&x) {}
  bool empty() const { return !m || m->empty(); } Mat getMat() const { return *m; } };
struct _OutputArray { Mat* m = nullptr; _OutputArray() {} _OutputArray(Mat& x) : m(&x) {}
  void release() const { *m = Mat(); } void create(int r, int c, int t) const { *m = Mat(r, c, t); } Mat getMat() const { Mat v; v.data = m->data; v.rows = m->rows; v.cols = m->cols; v.step = m->step; return v; } };
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
}  // namespace cv
