// TEST INFRASTRUCTURE -- NOT g2o (see ../types/types_six_dof_expmap.h).
#pragma once
#include "../types/types_six_dof_expmap.h"
