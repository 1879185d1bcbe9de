// TEST INFRASTRUCTURE ONLY (see orc_common.h).  The oracle of Optimizer::LocalInertialBA takes the flat graph
// of the C ABI (lia_graph_view, include/orb_b200.h) -- interface types only.
#pragma once
#include "../include/orb_b200.h"
