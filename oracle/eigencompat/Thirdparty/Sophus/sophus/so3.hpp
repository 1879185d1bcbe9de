// TEST INFRASTRUCTURE -- NOT Sophus: shadows the vendored header, which needs the real Eigen (see ../../../sophus/se3.hpp).
#pragma once
#include "../../../sophus/se3.hpp"
