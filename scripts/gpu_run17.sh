set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "bn_relu or native_vbm or cuda_graph or linear" > gpurun_out/pytest_bn3.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|assert " gpurun_out/pytest_bn3.log | head
python scripts/prof_bn.py
COINN_CONV1_IMPL=tc timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fwd_tc|apply4|conv1_wgrad_tc" -c 3 -o gpurun_out/prof_l1 -f python scripts/prof_bn.py prof > gpurun_out/ncu_l1.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_v7.log 2>&1; tail -1 gpurun_out/bench_v7.log | cut -c1-1000
timeout -s KILL 300 python bench.py --model fs --steps 300 --warmup 20 > gpurun_out/bench_fs.log 2>&1; tail -2 gpurun_out/bench_fs.log | cut -c1-900
