// TEST INFRASTRUCTURE -- NOT g2o: the kernel factory is not part of the compared code; registration expands to nothing.
#pragma once
#define G2O_REGISTER_ROBUST_KERNEL(name, classname)
