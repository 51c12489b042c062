set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "bn_relu or native_vbm or cuda_graph" > gpurun_out/pytest_bn2.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|assert " gpurun_out/pytest_bn2.log | head
python scripts/prof_bn.py
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_e2e2.log 2>&1; tail -1 gpurun_out/bench_e2e2.log | cut -c1-1000
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --graph 1 > gpurun_out/bench_graph.log 2>&1; tail -3 gpurun_out/bench_graph.log | cut -c1-1000
