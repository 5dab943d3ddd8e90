set -x
mkdir -p gpurun_out
export BATCH=1024 RUNS=3
TAG=split16 python scripts/profile_batch.py | tail -1 | tee gpurun_out/r2_ba_tune2.jsonl
for v in 8 32; do
  TAG=split$v G16_LIB=gnark_symmetric_crypto_b200/lib/variants/libg16b200_$v.so python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_ba_tune2.jsonl
done
TAG=split16_k2 G16_MSM_BA_K=2 python scripts/profile_batch.py | tail -1 | tee -a gpurun_out/r2_ba_tune2.jsonl
BATCH=512 RUNS=1 timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section SchedulerStats --section LaunchStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis \
   --clock-control none -k regex:msm_ba_ -c 9 -o gpurun_out/r2_ba_split -f python scripts/profile_batch.py > gpurun_out/r2_ba_ncu2.log 2>&1
ncu -i gpurun_out/r2_ba_split.ncu-rep --page details > gpurun_out/r2_ba_split_details.txt 2>&1
grep -E "msm_ba_|Duration|Registers Per|Achieved Occupancy|Issue Slots Busy|Executed Ipc Active|Warp Cycles Per Issued|DRAM Throughput|Avg. Active Threads" gpurun_out/r2_ba_split_details.txt | head -90
(time python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/r2_gputests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests3.log
tail -25 gpurun_out/r2_gputests3.log
