set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu_all.log | head -20
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_tcgen05conv.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_tcgen05conv.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_tcgen05conv.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench3.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_tcgen05conv.csv 24
