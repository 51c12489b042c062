set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "tcgen05_conv3d" > gpurun_out/pytest_conv2.log 2>&1; echo "conv rc=$?"
grep -E "passed|failed|Error|assert|FAILED" gpurun_out/pytest_conv2.log | head -40
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --skip-e2e > gpurun_out/bench_tma.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_tma.log | cut -c1-220
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_tma.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench4.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_tma.csv 16
