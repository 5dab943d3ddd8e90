# session 16: concurrent solves of the sub-batches (G16_SPLIT_SOLVE=2); the build with the SM count queried from the device
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep16.jsonl
TAG=default python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep16.jsonl
TAG=par_solve G16_SPLIT_SOLVE=2 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep16.jsonl
TAG=sub256 G16_SUBBATCH=256 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep16.jsonl
TAG=sub256_par G16_SUBBATCH=256 G16_SPLIT_SOLVE=2 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep16.jsonl
(time python -m pytest tests/test_gpu.py tests/test_gpu_round2.py tests/test_gpu_aes.py -m gpu -q -x) > gpurun_out/r2_gputests16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests16.log
tail -5 gpurun_out/r2_gputests16.log
