set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/r2_gputests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests2.log
tail -25 gpurun_out/r2_gputests2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ba.json 2> gpurun_out/r2_bench_ba.err; echo "bench rc=$?"
G16_MSM_BA=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-msm --no-strong > gpurun_out/r2_bench_noba.json 2> gpurun_out/r2_bench_noba.err
G16_MSM_BA_K=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-msm --no-strong > gpurun_out/r2_bench_ba_k2.json 2> gpurun_out/r2_bench_ba_k2.err
python - <<'PY'
import json
for n in ("ba","noba","ba_k2"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,2) for k,v in d["stages_ms_per_step"].items()}, d.get("msm_standalone"), d.get("msm_split"))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2_bench_{n}.err").read()[-1500:])
PY
