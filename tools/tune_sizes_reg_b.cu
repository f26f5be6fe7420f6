// register-direct input shape instances (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_reg_b() { add_reg_size<9>(); }
