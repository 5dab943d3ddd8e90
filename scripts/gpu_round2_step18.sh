# session 18: threshold of the direct leftovers (G16_MSM_BA_LEFT = t: remainders of <= t entries skip the batch-affine levels)
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep18.jsonl
for t in 3 4 5 6 7; do
  TAG=left_t$t G16_MSM_BA_LEFT=$t python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep18.jsonl
done
TAG=left_t7_c14 G16_MSM_BA_LEFT=7 G16_C_Z=14 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep18.jsonl
TAG=left_t5_c16 G16_MSM_BA_LEFT=5 G16_C_Z=16 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep18.jsonl
