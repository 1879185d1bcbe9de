// TEST INFRASTRUCTURE -- NOT g2o (see ../core/base_vertex.h): nothing of the type factory / macros is used by the compared code.
#pragma once
