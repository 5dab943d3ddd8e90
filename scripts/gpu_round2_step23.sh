# session 23: radix-4 register stage in the NTT pass (two stages per shared-memory round trip), 4 / 3 resident CTAs per SM
set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=4
rm -f gpurun_out/r2_sweep23.jsonl
TAG=radix2 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep23.jsonl
TAG=radix4_4cta_64regs G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_ntt_r4m4.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep23.jsonl
TAG=radix4_3cta_80regs G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_ntt_r4m3.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_sweep23.jsonl
