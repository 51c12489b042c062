set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -k "conv3d or vbm or conv1_fused" > gpurun_out/pytest_r33.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_r33.log | cut -c1-200; grep -E "FAILED|Error" gpurun_out/pytest_r33.log | head
timeout -s KILL 240 python scripts/prof_c1f.py 2>&1 | head -4
timeout -s KILL 600 python bench.py > gpurun_out/bench_r33.json 2> gpurun_out/bench_r33.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r33.json
