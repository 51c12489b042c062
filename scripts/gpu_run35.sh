set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r35.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r35.log | cut -c1-200
grep -E "FAILED|Error" gpurun_out/pytest_gpu_r35.log | head -20
timeout -s KILL 240 python scripts/prof_c1f.py 2>&1 | head -4
timeout -s KILL 600 python bench.py > gpurun_out/bench_r35.json 2> gpurun_out/bench_r35.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r35.json
timeout -s KILL 600 python bench.py --model fs > gpurun_out/bench_r35_fs.json 2> gpurun_out/bench_r35_fs.err; cut -c1-330 gpurun_out/bench_r35_fs.json
