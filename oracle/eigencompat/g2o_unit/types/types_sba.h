// TEST INFRASTRUCTURE -- NOT g2o (see ../core/base_vertex.h): the landmark vertex of types_sba.h, estimate only.
#pragma once
#include <g2o_base.h>
namespace g2o {
class VertexSBAPointXYZ : public BaseVertex<3, Eigen::Vector3d> {};
}  // namespace g2o
