#!/bin/bash
# Builds a variant of libse3icp_cuda.so with extra -D flags for A/B timing on the GPU box:
#   profiles/experiments/build_variant.sh NAME "-DKNN_INTERP=0 ..."   ->  se3-icp_b200/variants/libse3icp_NAME.so
# Select it with SE3ICP_LIB=se3-icp_b200/variants/libse3icp_NAME.so (se3-icp_b200/capi.py, development hook).
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
SRC=$ROOT/se3-icp_b200/csrc
OUT=$ROOT/se3-icp_b200/variants
TMP=$(mktemp -d)
mkdir -p "$OUT"
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in spatial_index se3_index knn_features shot_lrf nn_search optimise nccl_dyn capi eval; do
  $NVCC $ARCH -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr \
     $FLAGS -c $SRC/$f.cu -o $TMP/$f.o 2> $TMP/$f.log &
done
wait
grep -h -A1 "knn_features_kernel\|nn_search_kernel" $TMP/knn_features.log $TMP/nn_search.log | grep -v "^--" | grep "registers\|Compiling" | head -8
$NVCC $ARCH -shared -o $OUT/libse3icp_$NAME.so $TMP/*.o -lcudart -ldl
rm -rf $TMP
echo built $OUT/libse3icp_$NAME.so
