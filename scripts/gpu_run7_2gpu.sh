set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -6
timeout -s KILL 600 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/pytest_multigpu.log 2>&1; echo "multigpu rc=$?"
grep -E "passed|failed|Error|assert|FAILED" gpurun_out/pytest_multigpu.log | head -20
tail -30 gpurun_out/pytest_multigpu.log | cut -c1-300
NCCL_DEBUG=WARN timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench2 rc=$?"; tail -2 gpurun_out/bench_n2.log | cut -c1-600
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 5 --warmup 2 --impl reference > gpurun_out/bench_n2_ref.log 2>&1; echo "ref2 rc=$?"; tail -1 gpurun_out/bench_n2_ref.log | cut -c1-400
