# session 20: software prefetch of references / running products in the batch-affine addition; more solver chains
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep20.jsonl
TAG=prefetch python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep20.jsonl
TAG=no_prefetch G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_nopf.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep20.jsonl
TAG=prefetch_chains4 G16_SOLVE_CHAINS=4 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep20.jsonl
TAG=prefetch_chains3 G16_SOLVE_CHAINS=3 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep20.jsonl
TAG=prefetch_again python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep20.jsonl
