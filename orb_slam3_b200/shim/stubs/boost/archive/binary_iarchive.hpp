// SYNTAX-CHECK STAND-IN, not Boost.
#pragma once
#include "../serialization/serialization.hpp"
