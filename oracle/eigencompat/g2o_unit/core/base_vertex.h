// TEST INFRASTRUCTURE -- NOT g2o: what the vendored types_six_dof_expmap.{h,cpp} include by relative path, when they are
// compiled from a pipe with this directory tree as the working directory (oracle/Makefile, _ref/libref_g2o.so).
#pragma once
#include <g2o_base.h>
