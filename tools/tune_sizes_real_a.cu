// R2C / C2R shape instances (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_real_a() { add_real_size<5>(); add_real_size<6>(); add_real_size<7>(); add_real_size<8>(); }
