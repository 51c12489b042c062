set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r28.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_r28.log | cut -c1-200
grep -E "FAILED|Error" gpurun_out/pytest_gpu_r28.log | head -20
timeout -s KILL 600 python bench.py > gpurun_out/bench_r28.json 2> gpurun_out/bench_r28.err; echo "bench rc=$?"; cat gpurun_out/bench_r28.json | cut -c1-400
