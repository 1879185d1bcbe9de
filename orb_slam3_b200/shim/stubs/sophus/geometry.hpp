// SYNTAX-CHECK STAND-IN, not Sophus (see se3.hpp).
#pragma once
#include "se3.hpp"
