set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=3
TAG=noba G16_MSM_BA=0 python scripts/profile_batch.py | tail -1 | tee gpurun_out/r2_ba_tune.jsonl
TAG=default python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_ba_tune.jsonl
for v in 16x4 32x4 32x5 24x5 48x4 64x4; do
  TAG=$v G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_$v.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_ba_tune.jsonl
done
TAG=32x4_k2 G16_MSM_BA_K=2 G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_32x4.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_ba_tune.jsonl
BATCH=512 RUNS=1 timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section SchedulerStats --section LaunchStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis \
   --clock-control none -k regex:msm_ba_level_kernel -c 8 -o gpurun_out/r2_ba_levels -f python scripts/profile_batch.py > gpurun_out/r2_ba_ncu.log 2>&1
ncu -i gpurun_out/r2_ba_levels.ncu-rep --page details > gpurun_out/r2_ba_levels_details.txt 2>&1
grep -E "msm_ba_level_kernel|Duration|Registers|Achieved Occupancy|Issue Slots Busy|Executed Ipc|Warp Cycles Per Issued|Stall|L2 Hit|DRAM Throughput|Theoretical Occ|No Eligible|Eligible Warps" gpurun_out/r2_ba_levels_details.txt | head -120
