#!/bin/bash
# Times the reference's own CUDA backend (apps/qsim_base_cuda.cu recompiled for sm_100a,
# oracle/_ref/qsim_base_cuda_ref) on the same B200, same circuit, same fuser setting.
# usage: tools/ref_cuda_baseline.sh [f] -> prints the reference's "-v 2" timing lines
F=${1:-4}
R=oracle/_ref
for i in 1 2 3; do
  $R/qsim_base_cuda_ref -c $R/circuits/circuit_q30 -d 20 -f $F -v 2 2>&1 | grep -E "time|000" | tr '\n' ' '
  echo
done
