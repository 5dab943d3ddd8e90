# session 10: dedicated squaring (on / off variant), +-1 fast path of the digit kernel, Fp2-level out-of-line G2 products,
# block tree only for small levels; narrow windows for the wire queries as an env experiment
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_gputests10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests10.log
tail -8 gpurun_out/r2_gputests10.log
export BATCH=1024 RUNS=3
rm -f gpurun_out/r2_sweep10.jsonl
TAG=default python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep10.jsonl
[ -f gnark_symmetric_crypto_b200/lib/variants/libg16b200_nosqr.so ] && TAG=nosqr G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_nosqr.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep10.jsonl
TAG=c8_wire_queries G16_C_A=8 G16_C_B=8 G16_C_K=8 G16_C_B2=8 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep10.jsonl
TAG=nosplit_solve G16_SPLIT_SOLVE=0 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep10.jsonl
python scripts/latency.py 30 > gpurun_out/r2_latency10.log 2>&1; cp gpurun_out/latency.json gpurun_out/r2_latency10.json
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench10.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["stages_ms_per_step"])
print("msm", d["msm_standalone"].get("sizes")); print("single", d["single_request"]["ms_median"], d["single_request"]["stages_ms"])
PY
BATCH=1024 RUNS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches10.csv python scripts/profile_batch.py > gpurun_out/r2_launches10.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches10.csv > gpurun_out/r2_launches10_summary.txt; head -34 gpurun_out/r2_launches10_summary.txt
