# multi-GPU session: bench.py under torchrun at N GPUs (weak headline + strong_1024 + msm_split + library_multi_gpu), then the
# one-request API served from one process over all N GPUs (tools/serve_load).  usage: bash scripts/gpu_round2_multi.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_${N}gpu.json") if l.startswith("{")][-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "n_gpus", d["n_gpus"])
print("strong", d["strong_1024"]); print("split", d["msm_split"]); print("lib_multi", d["library_multi_gpu"])
PY
for callers in ${CALLERS:-1024 4096 8192}; do
  G16_DEVICES=all G16_BATCH_MAX=1024 gnark_symmetric_crypto_b200/lib/serve_load gnark_symmetric_crypto_b200/lib/libg16b200.so tests/golden/pk.chacha20 tests/golden/r1cs.chacha20 $callers 6 | tee -a gpurun_out/r2_serve_load_${N}gpu.jsonl
done
