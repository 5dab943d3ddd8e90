# session 22: last check of HEAD — GPU tests, smoke, bench line (both arms)
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2_gputests22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests22.log
tail -4 gpurun_out/r2_gputests22.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke22.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke22.log
(time python bench.py) > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_bench22.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench22.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["stages_ms_per_step"], d["gpu_launches"])
print({k: d["roofline"][k] for k in ("achieved", "peak", "frac", "traffic", "executed_products_per_addition", "executed_frac_of_modmul_peak")})
print(d["cpu_baseline"]); print({k: (round(v["value"], 1)) for k, v in d["aes"].items() if k.startswith("aes")}); print(d["verified"])
PY
