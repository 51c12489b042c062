set -x
mkdir -p gpurun_out
N=${1:-8}
timeout -s KILL 900 python -m pytest tests/test_multigpu.py -q > gpurun_out/pytest_multigpu_n$N.log 2>&1; echo "multigpu rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_multigpu_n$N.log | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout -s KILL 400 $TR --master-port 29901 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-900
timeout -s KILL 400 $TR --master-port 29902 bench.py --gpus $N --steps 8 --warmup 3 --impl reference > gpurun_out/bench_n${N}_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_n${N}_ref.log | cut -c1-400
timeout -s KILL 400 $TR --master-port 29903 bench.py --gpus $N --steps 100 --warmup 10 --transport nccl --skip-e2e > gpurun_out/bench_n${N}_nccl.log 2>&1; echo "nccl rc=$?"; tail -1 gpurun_out/bench_n${N}_nccl.log | cut -c1-300
timeout -s KILL 600 $TR --master-port 29904 bench_reduce.py --max-mb 256 --iters 20 > gpurun_out/reduce_sweep_n$N.log 2>&1; echo "sweep rc=$?"; tail -12 gpurun_out/reduce_sweep_n$N.log | cut -c1-260
timeout -s KILL 300 $TR --master-port 29905 bench.py --gpus $N --model fs --steps 500 --warmup 20 --skip-e2e > gpurun_out/bench_fs_n$N.log 2>&1; tail -1 gpurun_out/bench_fs_n$N.log | cut -c1-300
