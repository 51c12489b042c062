set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "conv1 or wgrad or cuda_graph" > gpurun_out/pytest_c1b.log 2>&1; echo "rc=$?"
grep -E "passed|failed|Error|assert|FAILED" gpurun_out/pytest_c1b.log | head -30
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --skip-e2e > gpurun_out/bench_c1tc2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c1tc2.log | cut -c1-220
COINN_CONV1_IMPL=cuda timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --skip-e2e > gpurun_out/bench_c1cuda.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c1cuda.log | cut -c1-220
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_c1tc2.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench7.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_c1tc2.csv 12
