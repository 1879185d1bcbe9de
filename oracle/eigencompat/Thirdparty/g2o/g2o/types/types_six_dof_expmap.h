// TEST INFRASTRUCTURE -- NOT g2o (the reference vendors g2o, but its headers need the real Eigen).  The vertex / edge
// base classes and SE3Quat as far as the reference's src/OptimizableTypes.cpp uses them, FUNCTIONAL: estimates,
// measurements, errors, Jacobian members, information, chi2 -- so that the reference's own computeError() /
// linearizeOplus() run inside oracle/_ref/libref_edges.so.  Follows Thirdparty/g2o/g2o/types/se3quat.h (map, *,
// normalizeRotation), core/base_vertex.h, base_unary_edge.h, base_binary_edge.h (member names only; no solver).
#pragma once
#include <g2o_base.h>
namespace g2o {
class VertexSE3Expmap : public BaseVertex<6, SE3Quat> {};
class VertexSBAPointXYZ : public BaseVertex<3, Eigen::Vector3d> {};
}  // namespace g2o
