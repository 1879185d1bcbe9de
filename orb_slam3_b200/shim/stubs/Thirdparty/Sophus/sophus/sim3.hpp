// SYNTAX-CHECK STAND-IN, not Sophus: shadows the vendored header, which needs Eigen (see ../../../sophus/se3.hpp).
#pragma once
#include "../../../sophus/se3.hpp"
