// shape instances for a group of sizes (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_d() { add_size<11>(); add_size<12>();  }
