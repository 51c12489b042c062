set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all2.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|assert " gpurun_out/pytest_gpu_all2.log | head
python scripts/prof_bn.py | head -8
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_v8.log 2>&1; tail -1 gpurun_out/bench_v8.log | cut -c1-1100
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --input-dtype fp32 --skip-e2e > gpurun_out/bench_v8_fp32in.log 2>&1; tail -1 gpurun_out/bench_v8_fp32in.log | cut -c1-300
timeout -s KILL 300 python bench.py --model fs --steps 500 --warmup 20 > gpurun_out/bench_fs.log 2>&1; tail -1 gpurun_out/bench_fs.log | cut -c1-900
timeout -s KILL 300 python bench.py --model fs --steps 500 --warmup 20 --graph 0 --skip-e2e > gpurun_out/bench_fs_eager.log 2>&1; tail -1 gpurun_out/bench_fs_eager.log | cut -c1-300
timeout -s KILL 300 python bench.py --model fs --steps 100 --warmup 5 --impl reference > gpurun_out/bench_fs_ref.log 2>&1; tail -1 gpurun_out/bench_fs_ref.log | cut -c1-300
