set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -x -q -k "linear or conv1 or bn_relu or native_vbm" > gpurun_out/pytest_vbm.log 2>&1; echo "vbm rc=$?"
tail -30 gpurun_out/pytest_vbm.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_native_cudnn.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_native_cudnn.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_native_cudnn.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench2.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_native_cudnn.csv
