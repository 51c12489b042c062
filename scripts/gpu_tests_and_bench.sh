set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r38.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r38.log | cut -c1-200
grep -E "FAILED|Error|assert" gpurun_out/pytest_gpu_r38.log | head -20 | cut -c1-220
timeout -s KILL 600 python bench.py > gpurun_out/bench_r38.json 2> gpurun_out/bench_r38.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r38.json; tail -3 gpurun_out/bench_r38.err
