#!/bin/bash
# one call: filter-kernel block shapes, source order on / off, reseed threshold (two interleaved passes)
run() { # label lib env...
  local label=$1 lib=$2; shift 2
  echo "[$label] $(env "$@" SE3ICP_LIB=$lib python profiles/experiments/ab_pair.py 20 2>&1 | tail -1)"
}
D=se3-icp_b200/libse3icp_cuda.so
V=se3-icp_b200/variants
for pass in 1 2; do
  run default $D X=1
  for v in f256x4 f256x5 f128x6 f128x8 f128x10; do run $v $V/libse3icp_$v.so X=1; done
  run noorder $D SE3ICP_SRC_ORDER=0
  run reseed0.02 $D SE3ICP_RESEED_THR=0.02
  run reseed0.15 $D SE3ICP_RESEED_THR=0.15
done
