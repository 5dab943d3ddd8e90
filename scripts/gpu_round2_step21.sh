# session 21: occupancy of the batch-affine addition (768 resident threads, 80 registers) and 64 pairs per thread
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep21.jsonl
TAG=default python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep21.jsonl
TAG=add_768_threads G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_t768.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep21.jsonl
TAG=pairs_per_thread_64 G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_m64.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep21.jsonl
