// SYNTAX-CHECK STAND-IN, not g2o (the reference vendors g2o, but its headers need Eigen, which this image lacks):
// the g2o types the reference's *headers* name in declarations.  The shims do not use g2o -- that is the point of them.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace g2o {
typedef Eigen::Matrix<double, 6, 1> Vector6d; typedef Eigen::Matrix<double, 7, 1> Vector7d;
typedef Eigen::Matrix<double, 3, 1> Vector3d; typedef Eigen::Matrix<double, 2, 1> Vector2d; typedef Eigen::Matrix<double, 3, 3> Matrix3d;
struct SE3Quat {
  SE3Quat() {} template <class... A> SE3Quat(const A&...) {}
  const Eigen::Vector3d& translation() const { static Eigen::Vector3d t; return t; } const Eigen::Quaterniond& rotation() const { static Eigen::Quaterniond q; return q; }
  SE3Quat inverse() const { return *this; } Eigen::Vector3d map(const Eigen::Vector3d&) const { return Eigen::Vector3d(); }
  Eigen::Matrix<double, 4, 4> to_homogeneous_matrix() const { return Eigen::Matrix<double, 4, 4>(); }
  SE3Quat operator*(const SE3Quat&) const { return *this; } static SE3Quat exp(const Vector6d&) { return SE3Quat(); } Vector6d log() const { return Vector6d(); }
};
struct Sim3 {
  Sim3() {} template <class... A> Sim3(const A&...) {}
  const Eigen::Vector3d& translation() const { static Eigen::Vector3d t; return t; } const Eigen::Quaterniond& rotation() const { static Eigen::Quaterniond q; return q; }
  double scale() const { return 1; } Sim3 inverse() const { return *this; } Eigen::Vector3d map(const Eigen::Vector3d&) const { return Eigen::Vector3d(); }
  Sim3 operator*(const Sim3&) const { return *this; } Vector7d log() const { return Vector7d(); }
};
namespace HyperGraph { struct Vertex { virtual ~Vertex() {} int id() const { return 0; } }; struct Edge { virtual ~Edge() {} }; }
namespace OptimizableGraph { struct Vertex : HyperGraph::Vertex {}; struct Edge : HyperGraph::Edge {}; }
template <int D, class T> struct BaseVertex : OptimizableGraph::Vertex { T _estimate; const T& estimate() const { return _estimate; } void setEstimate(const T& t) { _estimate = t; } };
template <int D, class E, class V1, class V2> struct BaseBinaryEdge : OptimizableGraph::Edge {};
template <int D, class E, class V1> struct BaseUnaryEdge : OptimizableGraph::Edge {};
template <int D, class E> struct BaseMultiEdge : OptimizableGraph::Edge {};
struct VertexSE3Expmap : BaseVertex<6, SE3Quat> {};
struct VertexSBAPointXYZ : BaseVertex<3, Eigen::Vector3d> {};
struct EdgeSE3ProjectXYZ : BaseBinaryEdge<2, Eigen::Vector2d, VertexSBAPointXYZ, VertexSE3Expmap> {};
struct EdgeStereoSE3ProjectXYZ : BaseBinaryEdge<3, Eigen::Vector3d, VertexSBAPointXYZ, VertexSE3Expmap> {};
struct SparseOptimizer {};
}  // namespace g2o
