# session 19: final records of the round (HEAD with direct leftovers): GPU tests, bench line, launch list, ncu --set full of the dominant kernel
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_gputests19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests19.log
tail -6 gpurun_out/r2_gputests19.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench19.err
BATCH=1024 RUNS=4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches19.csv python scripts/profile_batch.py > gpurun_out/r2_launches19.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches19.csv > gpurun_out/r2_launches19_summary.txt; head -12 gpurun_out/r2_launches19_summary.txt
# the level-0 batch-affine addition of the Z query in the 4th batch run (tables built): launches 0,1,2 are run 1 (levels 0,1,2), ... 9 = level 0 of run 4
BATCH=512 RUNS=4 timeout 900 ncu --set full --import-source on --clock-control none -k regex:msm_ba_add_kernel --launch-skip 9 -c 1 -o gpurun_out/r2_ba_add_full19 -f python scripts/profile_batch.py > gpurun_out/r2_ba_add_full19.log 2>&1
ncu -i gpurun_out/r2_ba_add_full19.ncu-rep --page details > gpurun_out/r2_ba_add_full19_details.txt 2>&1
ncu -i gpurun_out/r2_ba_add_full19.ncu-rep --page raw --csv > gpurun_out/r2_ba_add_full19_raw.csv 2>&1
grep -E "Duration|Registers Per|Grid Size" gpurun_out/r2_ba_add_full19_details.txt | head -5
# and the NTT pass + the group-sum accumulation, sections only
BATCH=512 RUNS=3 timeout 900 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --clock-control none -k regex:"ntt_pass_kernel|msm_accumulate_kernel|msm_tree_kernel|msm_digits_kernel|msm_ba_den_kernel" --launch-skip 40 -c 24 --page details --log-file gpurun_out/r2_sections19.txt python scripts/profile_batch.py > gpurun_out/r2_sections19.log 2>&1
grep -E "^  [a-z].*\(|Duration|DRAM Throughput|Registers Per|Achieved Occupancy|Executed Ipc Active" gpurun_out/r2_sections19.txt | head -80
