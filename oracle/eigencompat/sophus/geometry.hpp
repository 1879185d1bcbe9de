// TEST INFRASTRUCTURE -- NOT Sophus (see se3.hpp).
#pragma once
#include "se3.hpp"
