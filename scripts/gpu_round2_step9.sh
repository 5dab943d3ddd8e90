# session 9: block-parallel reduction tree (small MSMs / single request) on and off, full GPU tests, 1-GPU bench line
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_gputests9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests9.log
tail -6 gpurun_out/r2_gputests9.log
python scripts/latency.py 30 > gpurun_out/r2_latency_block.log 2>&1; cp gpurun_out/latency.json gpurun_out/r2_latency_tree_block.json
G16_MSM_TREE_BLOCK=0 python scripts/latency.py 30 > gpurun_out/r2_latency_serial.log 2>&1; cp gpurun_out/latency.json gpurun_out/r2_latency_tree_serial.json
python - <<'PY'
import json
for t in ("block", "serial"):
    d = json.load(open(f"gpurun_out/r2_latency_tree_{t}.json"))
    print(t, {k: v for k, v in d.items() if "ms" in k or "stages" in k})
PY
python scripts/sweep.py --max-log 20 > gpurun_out/r2_sweep_small_block.log 2>&1; cp gpurun_out/sweep.json gpurun_out/r2_sweep_small_block.json
G16_MSM_TREE_BLOCK=0 python scripts/sweep.py --max-log 20 > gpurun_out/r2_sweep_small_serial.log 2>&1; cp gpurun_out/sweep.json gpurun_out/r2_sweep_small_serial.json
tail -3 gpurun_out/r2_sweep_small_block.log gpurun_out/r2_sweep_small_serial.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench9.json
