set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=3
rm -f gpurun_out/r2_sweep3.jsonl
TAG=m32_adaptive python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
for sb in 256 1024; do TAG=sb$sb G16_SUBBATCH=$sb python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl; done
for c in 14 16; do TAG=cz$c G16_C_Z=$c python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl; done
TAG=k2 G16_MSM_BA_K=2 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
TAG=lmax32 G16_MSM_LMAX=32 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
TAG=lmax16 G16_MSM_LMAX=16 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
TAG=sideprio0 G16_SIDE_PRIO=0 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
TAG=c13_14 G16_C_A=14 G16_C_B=14 G16_C_K=14 G16_C_B2=14 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
TAG=c13_11 G16_C_A=11 G16_C_B=11 G16_C_K=11 G16_C_B2=11 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep3.jsonl
