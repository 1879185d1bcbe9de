// SYNTAX-CHECK STAND-IN, not Boost.
#pragma once
#include <string>
