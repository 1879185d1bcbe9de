// TEST INFRASTRUCTURE ONLY (oracle/_ref build): forwards to the minimal OpenCV stand-in.
#pragma once
#include "../opencv.hpp"
