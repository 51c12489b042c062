set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r30.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r30.log | cut -c1-200
grep -E "FAILED|Error" gpurun_out/pytest_gpu_r30.log | head -20
timeout -s KILL 600 python bench.py > gpurun_out/bench_r30.json 2> gpurun_out/bench_r30.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r30.json
timeout -s KILL 300 python scripts/prof_misc.py 2>&1 | tail -4
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 450 --csv --log-file gpurun_out/launches_v12.csv python bench.py --steps 4 --warmup 3 --graph 0 --skip-e2e > gpurun_out/ncu_v12.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_v12.csv 40 > gpurun_out/launches_v12.txt; head -32 gpurun_out/launches_v12.txt | cut -c1-150
