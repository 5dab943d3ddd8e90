# session 13: records of the final code of the round: GPU tests, 1-GPU bench line, launch list, ncu full of the dominant kernel
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_gputests13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests13.log
tail -6 gpurun_out/r2_gputests13.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench13.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench13_reference.json 2> gpurun_out/r2_bench13_reference.err; echo "reference rc=$?"
python scripts/latency.py 30 > gpurun_out/r2_latency13.log 2>&1; cp gpurun_out/latency.json gpurun_out/r2_latency13.json
python scripts/sweep.py --max-log 24 > gpurun_out/r2_sweep13.log 2>&1; cp gpurun_out/sweep.json gpurun_out/r2_sweep13.json
BATCH=1024 RUNS=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches13.csv python scripts/profile_batch.py > gpurun_out/r2_launches13.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches13.csv > gpurun_out/r2_launches13_summary.txt; head -45 gpurun_out/r2_launches13_summary.txt
# ncu full: the dominant kernel (level-0 batch-affine addition) of the 4th batch run (tables built, steady state)
BATCH=512 RUNS=4 timeout 900 ncu --set full --import-source on --clock-control none -k regex:msm_ba_add_kernel --launch-skip 9 -c 1 -o gpurun_out/r2_ba_add_full13 -f python scripts/profile_batch.py > gpurun_out/r2_ba_add_full13.log 2>&1
ncu -i gpurun_out/r2_ba_add_full13.ncu-rep --page details > gpurun_out/r2_ba_add_full13_details.txt 2>&1
ncu -i gpurun_out/r2_ba_add_full13.ncu-rep --page raw --csv > gpurun_out/r2_ba_add_full13_raw.csv 2>&1
grep -E "dram__bytes_read.sum,|dram__bytes_write.sum,|gpu__time_duration.sum|smsp__inst_executed.sum,|sm__inst_executed_pipe_fma|Kernel Name" gpurun_out/r2_ba_add_full13_raw.csv | head
