// TEST INFRASTRUCTURE -- NOT g2o: see core/optimization_algorithm_with_hessian.h (all stand-in interfaces live there).
#pragma once
#include "../core/optimization_algorithm_with_hessian.h"
