// TEST INFRASTRUCTURE -- intentionally empty: oracle/Makefile pipes the reference's own types_six_dof_expmap.h into the
// compiler ahead of types_six_dof_expmap.cpp; the self-include of the .cpp lands here.
