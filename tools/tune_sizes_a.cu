// shape instances for a group of sizes (split so the sweep compiles in parallel)
#include "tune_shapes.cuh"
void add_sizes_a() { add_size<5>(); add_size<6>();  }
