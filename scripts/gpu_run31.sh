set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -k "conv3d or vbm or conv1_fused" > gpurun_out/pytest_r31.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_r31.log | cut -c1-200
timeout -s KILL 600 python bench.py > gpurun_out/bench_r31.json 2> gpurun_out/bench_r31.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r31.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 450 --csv --log-file gpurun_out/launches_v13.csv python bench.py --steps 4 --warmup 3 --graph 0 --skip-e2e > gpurun_out/ncu_v13.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_v13.csv 40 > gpurun_out/launches_v13.txt; head -24 gpurun_out/launches_v13.txt | cut -c1-150
