// TEST INFRASTRUCTURE -- NOT Sophus (the reference vendors Sophus, but Sophus needs the real Eigen).  SE3 / SO3 over the
// functional stand-in of ../Eigen: construction from (quaternion | rotation matrix, translation), inverse(),
// rotationMatrix(), translation(), unit_quaternion(), products with SE3 and 3-vectors are FUNCTIONAL (what
// Frame::SetPose / UpdatePoseMatrices and ORBmatcher::SearchByProjection use), with the formulas of so3.hpp / se3.hpp
// (the rotation is a unit quaternion, points are rotated by Eigen's _transformVector); Sim3 and the Lie-algebra maps
// only type-check.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace Sophus {
template <class S> using Vector3 = Eigen::Matrix<S, 3, 1>;
template <class S> using Matrix3 = Eigen::Matrix<S, 3, 3>;
template <class S, int O = 0> struct SO3 {  // a unit quaternion, like Sophus (so3.hpp)
  typedef S Scalar;
  Eigen::Quaternion<S> q;
  SO3() {}
  SO3(const Eigen::Quaternion<S>& qq) : q(qq) {}
  template <class... A> SO3(const A&...) {}  // (rotation matrices ...: identity; not on a compared path)
  Eigen::Matrix<S, 3, 3> matrix() const { return q.toRotationMatrix(); }
  const Eigen::Quaternion<S>& unit_quaternion() const { return q; }
  SO3 inverse() const { return SO3(q.conjugate()); } Eigen::Matrix<S, 3, 1> log() const { return Eigen::Matrix<S, 3, 1>(); }
  template <class T> static SO3 exp(const T&) { return SO3(); }
  static Eigen::Matrix<S, 3, 3> hat(const Eigen::Matrix<S, 3, 1>& w) {
    Eigen::Matrix<S, 3, 3> W;
    W(0, 1) = -w[2]; W(0, 2) = w[1]; W(1, 0) = w[2]; W(1, 2) = -w[0]; W(2, 0) = -w[1]; W(2, 1) = w[0];
    return W;
  }
  template <class U> SO3<U> cast() const { return SO3<U>(q.template cast<U>()); }
  void normalize() {} S* data() { return nullptr; }
  SO3 operator*(const SO3& o) const { return SO3(q * o.q); }
  // so3.hpp: p' = unit_quaternion()._transformVector(p)
  template <int Op, int MR, int MC> Eigen::Matrix<S, 3, 1> operator*(const Eigen::Matrix<S, 3, 1, Op, MR, MC>& v) const { return q * v; }
};
template <class S, int O = 0> struct SE3 {
  typedef S Scalar;
  SO3<S> r;
  Eigen::Matrix<S, 3, 1> t;
  SE3() {}
  SE3(const Eigen::Quaternion<S>& q, const Eigen::Matrix<S, 3, 1>& tt) : r(q), t(tt) {}
  SE3(const SO3<S>& rr, const Eigen::Matrix<S, 3, 1>& tt) : r(rr), t(tt) {}
  template <class... A> SE3(const A&...) {}  // (rotation / 4x4 matrices ...: identity; not on a compared path)
  Eigen::Matrix<S, 3, 3> rotationMatrix() const { return r.matrix(); }
  const Eigen::Quaternion<S>& unit_quaternion() const { return r.unit_quaternion(); }
  Eigen::Matrix<S, 3, 1>& translation() { return t; }
  const Eigen::Matrix<S, 3, 1>& translation() const { return t; }
  const SO3<S>& so3() const { return r; }
  Eigen::Matrix<S, 4, 4> matrix() const { return Eigen::Matrix<S, 4, 4>(); } Eigen::Matrix<S, 3, 4> matrix3x4() const { return Eigen::Matrix<S, 3, 4>(); }
  // se3.hpp: inverse() = SE3(invR, invR * (translation() * Scalar(-1)))
  SE3 inverse() const { const SO3<S> invR = r.inverse(); return SE3(invR, invR * (t * S(-1))); }
  Eigen::Matrix<S, 6, 1> log() const { return Eigen::Matrix<S, 6, 1>(); }
  Eigen::Matrix<S, 6, 6> Adj() const { return Eigen::Matrix<S, 6, 6>(); } Eigen::Matrix<S, 7, 1> params() const { return Eigen::Matrix<S, 7, 1>(); }
  template <class T> static SE3 exp(const T&) { return SE3(); }
  template <class U> SE3<U> cast() const { return SE3<U>(r.template cast<U>(), t.template cast<U>()); }
  template <class T> void setQuaternion(const T&) {} template <class T> void setRotationMatrix(const T&) {} void normalize() {}
  S* data() { return nullptr; } const S* data() const { return nullptr; }
  // se3.hpp: SE3 * SE3 = (so3 * so3', t + so3 * t'),  SE3 * p = so3 * p + t
  SE3 operator*(const SE3& o) const { return SE3(r * o.r, t + r * o.t); }
  SE3& operator*=(const SE3& o) { *this = *this * o; return *this; }
  template <int Op, int MR, int MC> Eigen::Matrix<S, 3, 1> operator*(const Eigen::Matrix<S, 3, 1, Op, MR, MC>& v) const { return r * v + t; }
};
template <class S, int O = 0> struct RxSO3 {
  RxSO3() {} template <class... A> RxSO3(const A&...) {}
  S scale() const { return S(); } Eigen::Matrix<S, 3, 3> rotationMatrix() const { return Eigen::Matrix<S, 3, 3>(); }
  Eigen::Quaternion<S> quaternion() const { return Eigen::Quaternion<S>(); } Eigen::Matrix<S, 3, 3> matrix() const { return Eigen::Matrix<S, 3, 3>(); }
};
template <class S, int O = 0> struct Sim3 {
  typedef S Scalar;
  Sim3() {}
  template <class... A> Sim3(const A&...) {}
  S scale() const { return S(); } Eigen::Matrix<S, 3, 3> rotationMatrix() const { return Eigen::Matrix<S, 3, 3>(); }
  Eigen::Quaternion<S> quaternion() const { return Eigen::Quaternion<S>(); }
  Eigen::Matrix<S, 3, 1>& translation() { static Eigen::Matrix<S, 3, 1> t; return t; }
  const Eigen::Matrix<S, 3, 1>& translation() const { static Eigen::Matrix<S, 3, 1> t; return t; }
  RxSO3<S>& rxso3() { static RxSO3<S> r; return r; } const RxSO3<S>& rxso3() const { static RxSO3<S> r; return r; }
  Eigen::Matrix<S, 4, 4> matrix() const { return Eigen::Matrix<S, 4, 4>(); } Sim3 inverse() const { return *this; }
  Eigen::Matrix<S, 7, 1> log() const { return Eigen::Matrix<S, 7, 1>(); } template <class T> static Sim3 exp(const T&) { return Sim3(); }
  template <class U> Sim3<U> cast() const { return Sim3<U>(); } void setScale(S) {}
  Sim3 operator*(const Sim3&) const { return *this; }
  template <int R, int C, int Op, int MR, int MC> Eigen::Matrix<S, 3, 1> operator*(const Eigen::Matrix<S, R, C, Op, MR, MC>&) const { return Eigen::Matrix<S, 3, 1>(); }
};
typedef SO3<float> SO3f; typedef SO3<double> SO3d; typedef SE3<float> SE3f; typedef SE3<double> SE3d;
typedef Sim3<float> Sim3f; typedef Sim3<double> Sim3d; typedef RxSO3<float> RxSO3f; typedef RxSO3<double> RxSO3d;
template <class T> std::ostream& operator<<(std::ostream& o, const SE3<T>&) { return o; }
}  // namespace Sophus
