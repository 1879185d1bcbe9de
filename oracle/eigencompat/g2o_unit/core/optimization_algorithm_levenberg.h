// TEST INFRASTRUCTURE -- intentionally empty: oracle/Makefile pipes the reference's own optimization_algorithm_levenberg.h
// into the compiler ahead of optimization_algorithm_levenberg.cpp; the self-include of the .cpp lands here.
