set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_gputests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests4.log
tail -6 gpurun_out/r2_gputests4.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench2.err
BATCH=512 RUNS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:msm_ba_add_kernel -c 1 -o gpurun_out/r2_ba_add_full -f python scripts/profile_batch.py > gpurun_out/r2_ba_add_full.log 2>&1
ncu -i gpurun_out/r2_ba_add_full.ncu-rep --page details > gpurun_out/r2_ba_add_full_details.txt 2>&1
ncu -i gpurun_out/r2_ba_add_full.ncu-rep --page raw --csv > gpurun_out/r2_ba_add_full_raw.csv 2>&1
grep -E "dram__bytes_read.sum,|dram__bytes_write.sum,|gpu__time_duration.sum|smsp__inst_executed.sum,|sm__inst_executed_pipe_fma" gpurun_out/r2_ba_add_full_raw.csv | head
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench2.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["stages_ms_per_step"])
print(json.dumps(d["roofline"], indent=0)[:1500])
print(d["strong_1024"], d["msm_split"], d["msm_standalone"])
PY
