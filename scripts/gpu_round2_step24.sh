# session 24: two-lane schedule as the default for circuits without a commitment
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2_gputests24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests24.log
tail -5 gpurun_out/r2_gputests24.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench24.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench24.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["ms_per_step"], d["schedule"], d["stages_ms_per_step"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "executed_frac_of_modmul_peak")}, d["strong_1024"]["value"], d["single_request"]["ms_median"], {k: round(v["value"], 1) for k, v in d["aes"].items() if k.startswith("aes")}, d["verified"]["device_ms"])
PY
python scripts/latency.py 20 > gpurun_out/r2_latency24.log 2>&1; python - <<'PY'
import json
l = json.load(open("gpurun_out/latency.json"))
print(l["prove_json_ms"], [(b, round(l[b]["ms_median"], 2)) for b in ("batch64", "batch128", "batch256")])
PY
