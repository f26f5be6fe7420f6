// R2C / C2R shape instances (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_real_c() { add_real_size<11>(); add_real_size<12>(); }
