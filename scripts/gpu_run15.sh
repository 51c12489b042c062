set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "conv1" > gpurun_out/pytest_c1c.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_c1c.log | head
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_e2e.log 2>&1; tail -1 gpurun_out/bench_e2e.log | cut -c1-1200
COINN_CONV1_IMPL=tc timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --skip-e2e > gpurun_out/bench_c1tc3.log 2>&1; tail -1 gpurun_out/bench_c1tc3.log | cut -c1-200
COINN_CONV1_IMPL=tc timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_c1tc3.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench10.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_c1tc3.csv 8
