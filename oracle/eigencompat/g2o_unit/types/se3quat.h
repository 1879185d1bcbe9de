// TEST INFRASTRUCTURE -- NOT g2o: where the piped types_six_dof_expmap.h's `#include "se3quat.h"` lands (oracle/Makefile,
// working directory g2o_unit/types).  Forwards to the REFERENCE's own header by the absolute path the recipe passes in
// (-DG2O_REF_TYPES_DIR=..., together with -DG2O_BASE_REAL_SE3QUAT which drops g2o_base.h's own SE3Quat), so SE3Quat -- exp, log, map, operator*, inverse, normalizeRotation -- and skew / deltaR are
// the reference's code, over the functional Eigen stand-in; the vertex / edge bases stay those of g2o_base.h.
#pragma once
#define G2O_STR2(x) #x
#define G2O_STR(x) G2O_STR2(x)
#include G2O_STR(G2O_REF_TYPES_DIR/se3quat.h)
#include <g2o_base.h>
