// shape instances for a group of sizes (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_b() { add_size<7>(); add_size<8>();  }
