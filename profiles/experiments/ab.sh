#!/bin/bash
# A/B timing of library variants on the GPU box: profiles/experiments/ab.sh script.py variant...
S=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then L=se3-icp_b200/libse3icp_cuda.so; else L=se3-icp_b200/variants/libse3icp_$v.so; fi
  for rep in 1 2; do echo "[$v] $(SE3ICP_LIB=$L python $S 2>&1 | tr '\n' ' ')"; done
done
