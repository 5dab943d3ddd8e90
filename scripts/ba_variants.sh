#!/bin/bash
# Builds experimental variants of the batch-affine level kernels (pairs per thread M) as separate
# shared libraries under gnark_symmetric_crypto_b200/lib/variants/ for on-GPU tuning (scripts/profile_batch.py with G16_LIB).
set -e
cd "$(dirname "$0")/../gnark_symmetric_crypto_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p ../lib/variants _obj
for v in "$@"; do
  M=$v
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
        -DG16_BA_M=$M -c k_msm_ba.cu -o _obj/k_msm_ba_${v}.o
  OBJS=$(ls _obj/*.o | grep -v "k_msm_ba")
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libg16b200_${v}.so $OBJS _obj/k_msm_ba_${v}.o -lcudart_static -lpthread -ldl -lrt
  echo built $v
done
