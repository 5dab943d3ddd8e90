set -x
mkdir -p gpurun_out
BATCH=1024 RUNS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_ba.csv python scripts/profile_batch.py > gpurun_out/r2_launches_ba.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_ba.csv | tee gpurun_out/r2_launches_ba_summary.txt
