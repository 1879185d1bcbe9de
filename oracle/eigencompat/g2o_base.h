// TEST INFRASTRUCTURE -- NOT g2o (the reference vendors g2o, but its core headers need the real Eigen).  SE3Quat and the
// vertex / edge base classes as far as the reference's src/OptimizableTypes.cpp and the vendored
// Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp} use them, FUNCTIONAL: estimates, measurements, errors, Jacobian
// members, information, chi2 -- so that the reference's own computeError() / linearizeOplus() run inside oracle/_ref/.
// Follows Thirdparty/g2o/g2o/types/se3quat.h (map, *, inverse, normalizeRotation), core/base_vertex.h, base_unary_edge.h,
// base_binary_edge.h (member names only; no solver).
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <iostream>
#include <set>
#include <vector>
namespace g2o {
typedef Eigen::Matrix<double, 6, 1> Vector6d; typedef Eigen::Matrix<double, 7, 1> Vector7d;
typedef Eigen::Matrix<double, 3, 1> Vector3d; typedef Eigen::Matrix<double, 2, 1> Vector2d; typedef Eigen::Matrix<double, 3, 3> Matrix3d;
using Eigen::Quaterniond;
#ifndef G2O_BASE_REAL_SE3QUAT   // (oracle/Makefile, libref_g2o.so: the reference's own se3quat.h is in use instead)
class SE3Quat {
 public:
  SE3Quat() {}
  SE3Quat(const Eigen::Quaterniond& q, const Eigen::Vector3d& t) : _r(q), _t(t) { normalizeRotation(); }
  template <class... A> SE3Quat(const A&...) {}  // (rotation matrix, t), Vector6d / Vector7d: identity (not on a compared path)
  const Eigen::Vector3d& translation() const { return _t; } const Eigen::Quaterniond& rotation() const { return _r; }
  void setTranslation(const Eigen::Vector3d& t) { _t = t; } void setRotation(const Eigen::Quaterniond& r) { _r = r; }
  Eigen::Vector3d map(const Eigen::Vector3d& xyz) const { return _r * xyz + _t; }  // se3quat.h: _r*xyz + _t
  SE3Quat operator*(const SE3Quat& tr2) const {  // se3quat.h:104-110
    SE3Quat result(*this);
    result._t += _r * tr2._t;
    result._r *= tr2._r;
    result.normalizeRotation();
    return result;
  }
  SE3Quat inverse() const { SE3Quat ret; ret._r = _r.conjugate(); ret._t = ret._r * (_t * -1.); return ret; }
  void normalizeRotation() { if (_r.w() < 0) { _r.x() *= -1; _r.y() *= -1; _r.z() *= -1; _r.w() *= -1; } _r.normalize(); }
  Eigen::Matrix<double, 4, 4> to_homogeneous_matrix() const { return Eigen::Matrix<double, 4, 4>(); }
  static SE3Quat exp(const Vector6d&) { return SE3Quat(); } Vector6d log() const { return Vector6d(); }
  double operator[](int i) const { return i < 3 ? _t[i] : (i == 3 ? _r.x() : i == 4 ? _r.y() : i == 5 ? _r.z() : _r.w()); }
  void fromVector(const Vector7d& v) { _r = Eigen::Quaterniond(v[6], v[3], v[4], v[5]); _t = Eigen::Vector3d(v[0], v[1], v[2]); }
 protected:
  Eigen::Quaterniond _r;
  Eigen::Vector3d _t;
};
#endif
class Sim3 {  // type-check only
 public:
  Sim3() {} template <class... A> Sim3(const A&...) {}
  const Eigen::Vector3d& translation() const { static Eigen::Vector3d t; return t; } const Eigen::Quaterniond& rotation() const { static Eigen::Quaterniond q; return q; }
  double scale() const { return 1; } Sim3 inverse() const { return *this; } Eigen::Vector3d map(const Eigen::Vector3d& x) const { return x; }
  Sim3 operator*(const Sim3&) const { return *this; } Vector7d log() const { return Vector7d(); }
};
class HyperGraph {
 public:
  class Vertex { public: virtual ~Vertex() {} int id() const { return _id; } void setId(int i) { _id = i; } protected: int _id = 0; };
  class Edge {
   public:
    virtual ~Edge() {}
    void resize(size_t n) { _vertices.resize(n, nullptr); }
    void setVertex(size_t i, Vertex* v) { _vertices[i] = v; }
    Vertex* vertex(size_t i) { return _vertices[i]; }
   protected:
    std::vector<Vertex*> _vertices;
  };
};
class RobustKernel;
class OptimizableGraph : public HyperGraph {
 public:
  typedef std::set<HyperGraph::Vertex*> VertexSet;
  class Vertex : public HyperGraph::Vertex {
   public:
    virtual bool read(std::istream&) { return true; } virtual bool write(std::ostream&) const { return true; }
    virtual void setToOriginImpl() {} virtual void oplusImpl(const double*) {}
    void oplus(const double* v) { oplusImpl(v); updateCache(); }   // optimizable_graph.h:296-300
    void setFixed(bool f) { _fixed = f; } bool fixed() const { return _fixed; } void setMarginalized(bool m) { _marg = m; }
    void updateCache() {}
   protected:
    bool _fixed = false, _marg = false, _marginalized = false;
  };
  class Edge : public HyperGraph::Edge {
   public:
    virtual bool read(std::istream&) = 0; virtual bool write(std::ostream&) const = 0;
    virtual void computeError() = 0; virtual void linearizeOplus() = 0;
    void setRobustKernel(RobustKernel* k) { _rk = k; } RobustKernel* robustKernel() const { return _rk; }
    void setLevel(int l) { _level = l; } int level() const { return _level; }
   protected:
    RobustKernel* _rk = nullptr; int _level = 0;
  };
};
template <int D, class T> class BaseVertex : public OptimizableGraph::Vertex {
 public:
  static const int Dimension = D;
  typedef T EstimateType;
  const T& estimate() const { return _estimate; } void setEstimate(const T& t) { _estimate = t; }
  // base_vertex.h: the vertex's block of the Hessian and of the right-hand side (there: Maps into the solver's memory)
  Eigen::Matrix<double, D, 1>& b() { return _b; } Eigen::Matrix<double, D, D>& A() { return _hessianBlock; }
  void clearQuadraticForm() { _b.setZero(); _hessianBlock.setZero(); }
 protected:
  T _estimate;
  Eigen::Matrix<double, D, 1> _b; Eigen::Matrix<double, D, D> _hessianBlock;
};
template <int D, class E> class BaseEdge : public OptimizableGraph::Edge {
 public:
  typedef E Measurement;
  typedef Eigen::Matrix<double, D, 1> ErrorVector;
  typedef Eigen::Matrix<double, D, D> InformationType;
  BaseEdge() { _information.setIdentity(); }
  const E& measurement() const { return _measurement; } void setMeasurement(const E& m) { _measurement = m; }
  const InformationType& information() const { return _information; } InformationType& information() { return _information; }
  void setInformation(const InformationType& i) { _information = i; }
  const ErrorVector& error() const { return _error; }
  double chi2() const { return _error.dot(_information * _error); }
  // base_edge.h:102-109 (the second-order term is commented out there as well)
  InformationType robustInformation(const Eigen::Vector3d& rho) { InformationType result = rho[1] * _information; return result; }
 protected:
  E _measurement; InformationType _information; ErrorVector _error;
};
template <int D, class E, class VertexXi> class BaseUnaryEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
  typedef typename BaseEdge<D, E>::InformationType InformationType; typedef typename BaseEdge<D, E>::ErrorVector ErrorVector;
  BaseUnaryEdge() { this->resize(1); }
  const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
  void constructQuadraticForm();   // defined by the reference's base_unary_edge.hpp where oracle/Makefile pipes it in
 protected:
  using BaseEdge<D, E>::_measurement; using BaseEdge<D, E>::_information; using BaseEdge<D, E>::_error; using HyperGraph::Edge::_vertices;
  JacobianXiOplusType _jacobianOplusXi;
};
template <int D, class E, class VertexXi, class VertexXj> class BaseBinaryEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
  typedef Eigen::Matrix<double, D, VertexXj::Dimension> JacobianXjOplusType;
  typedef typename BaseEdge<D, E>::InformationType InformationType; typedef typename BaseEdge<D, E>::ErrorVector ErrorVector;
  BaseBinaryEdge() { this->resize(2); }
  const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
  const JacobianXjOplusType& jacobianOplusXj() const { return _jacobianOplusXj; }
  // base_binary_edge.h: the off-diagonal block this edge contributes to (there: Maps into the solver's memory)
  typedef Eigen::Matrix<double, VertexXi::Dimension, VertexXj::Dimension> HessianBlockType;
  typedef Eigen::Matrix<double, VertexXj::Dimension, VertexXi::Dimension> HessianBlockTransposedType;
  void constructQuadraticForm();   // defined by the reference's base_binary_edge.hpp where oracle/Makefile pipes it in
  const HessianBlockType& hessianBlock() const { return _hessian; }
 protected:
  using BaseEdge<D, E>::_measurement; using BaseEdge<D, E>::_information; using BaseEdge<D, E>::_error; using HyperGraph::Edge::_vertices;
  JacobianXiOplusType _jacobianOplusXi; JacobianXjOplusType _jacobianOplusXj;
  bool _hessianRowMajor = false; HessianBlockType _hessian; HessianBlockTransposedType _hessianTransposed;
};
template <int D, class E> class BaseMultiEdge : public BaseEdge<D, E> {  // (G2oTypes.h's inertial edges: members only)
 public:
  typedef Eigen::Matrix<double, D, Eigen::Dynamic> JacobianType;
 protected:
  using BaseEdge<D, E>::_measurement; using BaseEdge<D, E>::_information; using BaseEdge<D, E>::_error; using HyperGraph::Edge::_vertices;
  std::vector<JacobianType> _jacobianOplus;
};
}  // namespace g2o
