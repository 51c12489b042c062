set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r22.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_r22.log
grep -E "FAILED|Error" gpurun_out/pytest_gpu_r22.log | head -20
timeout -s KILL 300 python scripts/prof_misc.py > gpurun_out/prof_misc_r22.log 2>&1; tail -15 gpurun_out/prof_misc_r22.log
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke_r22.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r22.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_r22.json 2> gpurun_out/bench_r22.err; echo "bench rc=$?"; cat gpurun_out/bench_r22.json
timeout -s KILL 600 python bench.py --overlap 1 > gpurun_out/bench_r22_overlap.json 2> gpurun_out/bench_r22_overlap.err; echo "bench rc=$?"; cat gpurun_out/bench_r22_overlap.json
timeout -s KILL 600 python bench.py --model fs > gpurun_out/bench_r22_fs.json 2> gpurun_out/bench_r22_fs.err; cat gpurun_out/bench_r22_fs.json
