set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "tcgen05_conv3d" > gpurun_out/pytest_wh.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_wh.log | head -30
python scripts/prof_conv.py time
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --skip-e2e > gpurun_out/bench_wh.log 2>&1; tail -1 gpurun_out/bench_wh.log | cut -c1-200
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_wh.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench9.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_wh.csv 14
