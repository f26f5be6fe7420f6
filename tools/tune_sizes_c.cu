// shape instances for a group of sizes (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_c() { add_size<9>(); add_size<10>();  }
