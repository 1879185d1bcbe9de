// SYNTAX-CHECK STAND-IN, not Sophus (the reference vendors Sophus, but Sophus needs Eigen, which this image lacks):
// the members of SO3 / SE3 / Sim3 that the reference's headers and this repo's shims name, over the Eigen stand-in.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace Sophus {
template <class S> using Vector3 = Eigen::Matrix<S, 3, 1>;
template <class S> using Matrix3 = Eigen::Matrix<S, 3, 3>;
template <class S, int O = 0> struct SO3 {
  typedef S Scalar;
  SO3() {}
  template <class... A> SO3(const A&...) {}
  Eigen::Matrix<S, 3, 3> matrix() const { return Eigen::Matrix<S, 3, 3>(); }
  Eigen::Quaternion<S> unit_quaternion() const { return Eigen::Quaternion<S>(); }
  SO3 inverse() const { return *this; } Eigen::Matrix<S, 3, 1> log() const { return Eigen::Matrix<S, 3, 1>(); }
  template <class T> static SO3 exp(const T&) { return SO3(); }
  template <class T> static Eigen::Matrix<S, 3, 3> hat(const T&) { return Eigen::Matrix<S, 3, 3>(); }
  template <class U> SO3<U> cast() const { return SO3<U>(); }
  void normalize() {} S* data() { return nullptr; }
  SO3 operator*(const SO3&) const { return *this; }
  template <int R, int C, int Op, int MR, int MC> Eigen::Matrix<S, 3, 1> operator*(const Eigen::Matrix<S, R, C, Op, MR, MC>&) const { return Eigen::Matrix<S, 3, 1>(); }
};
template <class S, int O = 0> struct SE3 {
  typedef S Scalar;
  SE3() {}
  template <class... A> SE3(const A&...) {}
  Eigen::Matrix<S, 3, 3> rotationMatrix() const { return Eigen::Matrix<S, 3, 3>(); }
  Eigen::Quaternion<S> unit_quaternion() const { return Eigen::Quaternion<S>(); }
  Eigen::Matrix<S, 3, 1>& translation() { static Eigen::Matrix<S, 3, 1> t; return t; }
  const Eigen::Matrix<S, 3, 1>& translation() const { static Eigen::Matrix<S, 3, 1> t; return t; }
  SO3<S>& so3() { static SO3<S> r; return r; } const SO3<S>& so3() const { static SO3<S> r; return r; }
  Eigen::Matrix<S, 4, 4> matrix() const { return Eigen::Matrix<S, 4, 4>(); } Eigen::Matrix<S, 3, 4> matrix3x4() const { return Eigen::Matrix<S, 3, 4>(); }
  SE3 inverse() const { return *this; } Eigen::Matrix<S, 6, 1> log() const { return Eigen::Matrix<S, 6, 1>(); }
  Eigen::Matrix<S, 6, 6> Adj() const { return Eigen::Matrix<S, 6, 6>(); } Eigen::Matrix<S, 7, 1> params() const { return Eigen::Matrix<S, 7, 1>(); }
  template <class T> static SE3 exp(const T&) { return SE3(); }
  template <class U> SE3<U> cast() const { return SE3<U>(); }
  template <class T> void setQuaternion(const T&) {} template <class T> void setRotationMatrix(const T&) {} void normalize() {}
  S* data() { return nullptr; } const S* data() const { return nullptr; }
  SE3 operator*(const SE3&) const { return *this; } SE3& operator*=(const SE3&) { return *this; }
  template <int R, int C, int Op, int MR, int MC> Eigen::Matrix<S, 3, 1> operator*(const Eigen::Matrix<S, R, C, Op, MR, MC>&) const { return Eigen::Matrix<S, 3, 1>(); }
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
