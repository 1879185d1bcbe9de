// SYNTAX-CHECK STAND-IN, not Boost (see serialization.hpp).
#pragma once
#include "serialization.hpp"
