# session 12: two-level inversion of the batch-affine levels; split solve again now that the side stream is light
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep12.jsonl
TAG=inv2 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep12.jsonl
TAG=inv1 G16_BA_INV2=0 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep12.jsonl
TAG=inv2_splitsolve G16_SPLIT_SOLVE=1 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep12.jsonl
TAG=inv2_min16 G16_BA_INV2_MIN=65536 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep12.jsonl
(time python -m pytest tests/test_gpu.py tests/test_gpu_round2.py -m gpu -q -x -k "kat or batch or msm_paths or bit_wire") > gpurun_out/r2_gputests12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests12.log
tail -6 gpurun_out/r2_gputests12.log
