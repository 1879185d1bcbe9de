// SYNTAX-CHECK STAND-IN, not OpenCV (see ../opencv.hpp).
#pragma once
#include "../opencv.hpp"
