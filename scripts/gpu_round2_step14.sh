# session 14: ternary groups ({0, 1, -1} wires) in the combination tables, no batch-affine levels for the 0/+-1 queries
set -x
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "bit_wire") > gpurun_out/r2_gputests14a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests14a.log
tail -15 gpurun_out/r2_gputests14a.log
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep14.jsonl
TAG=bitq python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep14.jsonl
TAG=nobitq G16_BITQ=0 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep14.jsonl
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_gputests14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests14.log
tail -8 gpurun_out/r2_gputests14.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench14.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench14.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["stages_ms_per_step"])
print("aes", d["aes"])
PY
BATCH=1024 RUNS=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches14.csv python scripts/profile_batch.py > gpurun_out/r2_launches14.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches14.csv > gpurun_out/r2_launches14_summary.txt; head -40 gpurun_out/r2_launches14_summary.txt
