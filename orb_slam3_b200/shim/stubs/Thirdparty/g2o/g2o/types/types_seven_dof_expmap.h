// SYNTAX-CHECK STAND-IN, not g2o (see ../types/types_six_dof_expmap.h).
#pragma once
#include "../types/types_six_dof_expmap.h"
