set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=3
rm -f gpurun_out/r2_sweep4.jsonl
TAG=rowsort python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep4.jsonl
TAG=norowsort G16_MSM_ROWSORT=0 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep4.jsonl
TAG=rowsort_sb1024 G16_SUBBATCH=1024 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep4.jsonl
(time python -m pytest tests/test_gpu_round2.py tests/test_gpu.py -m gpu -x -q) > gpurun_out/r2_gputests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests5.log
tail -6 gpurun_out/r2_gputests5.log
BATCH=1024 RUNS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_rs.csv python scripts/profile_batch.py > gpurun_out/r2_launches_rs.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_rs.csv | head -14
