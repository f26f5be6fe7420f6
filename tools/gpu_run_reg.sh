#!/bin/bash
mkdir -p gpurun_out
for cv in -1 100 50 30; do
  TUNE_CARVEOUT=$cv timeout 400 tools/tune_reg 29 5 0 2 > gpurun_out/tune_reg_cv$cv.csv 2> gpurun_out/tune_reg_cv$cv.err; echo "cv=$cv rc=$? rows=$(wc -l < gpurun_out/tune_reg_cv$cv.csv)"
done
