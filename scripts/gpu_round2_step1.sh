set -x
mkdir -p gpurun_out
nvidia-smi -L
(time python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -30 gpurun_out/r2_gputests.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
(time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py) > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
tail -15 gpurun_out/r2_memcheck.log
