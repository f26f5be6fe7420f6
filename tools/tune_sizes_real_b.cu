// R2C / C2R shape instances (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_real_b() { add_real_size<9>(); add_real_size<10>(); }
