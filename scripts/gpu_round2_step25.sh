# session 25: last check of HEAD (two lanes, balanced sub-batches) — GPU tests, smoke, default bench run, 768-proof batch
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2_gputests25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests25.log
tail -4 gpurun_out/r2_gputests25.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke25.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke25.log
(time python bench.py) > gpurun_out/r2_bench25.json 2> gpurun_out/r2_bench25.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_bench25.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench25.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["stages_ms_per_step"], d["gpu_launches"])
print({k: d["roofline"][k] for k in ("achieved", "peak", "frac", "traffic", "executed_products_per_addition", "executed_frac_of_modmul_peak")})
print(d["cpu_baseline"]); print({k: (round(v["value"], 1)) for k, v in d["aes"].items() if k.startswith("aes")}); print(d["verified"])
PY
BATCH=768 RUNS=4 TAG=batch768 python scripts/profile_batch.py | tail -1
BATCH=1536 RUNS=4 TAG=batch1536 python scripts/profile_batch.py | tail -1
