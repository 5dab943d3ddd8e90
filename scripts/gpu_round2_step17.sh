# session 17: direct leftovers of the batch-affine runs (remainders of <= 5 entries skip the levels), concurrent solves as the default
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep17.jsonl
TAG=leftovers python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep17.jsonl
TAG=plain_padding G16_MSM_BA_LEFT=0 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep17.jsonl
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r2_gputests17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests17.log
tail -5 gpurun_out/r2_gputests17.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench17.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench17.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["stages_ms_per_step"])
r = d["roofline"]; print({k: r[k] for k in ("achieved", "frac", "executed_products_per_addition", "executed_frac_of_modmul_peak")})
print([(x["log2n"], round(x["one_shot"]["ms"], 2), round(x["fixed_base"]["ms"], 2)) for x in d["msm_standalone"]["sizes"]], d["msm_split"])
PY
BATCH=1024 RUNS=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches17.csv python scripts/profile_batch.py > gpurun_out/r2_launches17.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches17.csv > gpurun_out/r2_launches17_summary.txt; head -16 gpurun_out/r2_launches17_summary.txt; tail -1 gpurun_out/r2_launches17_summary.txt
