// dual-lane shape instances (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_dual_c() { add_dual_size<10>(); }
