#!/bin/bash
# one call: L1 prefetch hints in the 1-NN traversals (NN_PREFETCH = 1 leaf rows, 2 child boxes, 3 both), two interleaved passes
D=se3-icp_b200/libse3icp_cuda.so
V=se3-icp_b200/variants
lib() { if [ "$1" = default ]; then echo $D; else echo $V/libse3icp_$1.so; fi; }
for pass in 1 2; do
  for v in default pf1 pf2 pf3; do
    echo "[$v] $(SE3ICP_LIB=$(lib $v) python profiles/experiments/ab_pair.py 20 2>&1 | tail -1)"
  done
done
for v in default pf1 pf3 default; do
  echo "[$v] $(SE3ICP_LIB=$(lib $v) python profiles/experiments/quick_tput.py 2>&1 | tail -1)"
done
