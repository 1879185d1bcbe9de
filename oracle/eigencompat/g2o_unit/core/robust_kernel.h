// TEST INFRASTRUCTURE -- intentionally empty: oracle/Makefile pipes the reference's own robust_kernel.h, robust_kernel.cpp,
// robust_kernel_impl.h and robust_kernel_impl.cpp into the compiler in that order; their includes of each other land here.
