#!/bin/bash
# profiles/experiments/ab_env.sh script.py variant... ; honours SE3ICP_SE3_ORDER from the environment
S=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then L=se3-icp_b200/libse3icp_cuda.so; else L=se3-icp_b200/variants/libse3icp_$v.so; fi
  for rep in 1 2; do echo "[$v ${SE3ICP_SE3_ORDER:-kd}] $(SE3ICP_LIB=$L python $S 2>&1 | tr '\n' ' ')"; done
done
