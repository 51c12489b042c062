set -x
mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 900 python -m pytest tests/test_multigpu.py -q > gpurun_out/pytest_multigpu_v2_n$N.log 2>&1; echo "multigpu rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_multigpu_v2_n$N.log | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29901 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_v2_n$N.json 2> gpurun_out/bench_v2_n$N.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_v2_n$N.json
timeout -s KILL 400 $TR --master-port 29902 bench.py --gpus $N --steps 200 --warmup 10 --overlap 1 --skip-e2e > gpurun_out/bench_v2_n${N}_overlap.json 2> gpurun_out/bench_v2_n${N}_overlap.err; echo "overlap rc=$?"; cut -c1-260 gpurun_out/bench_v2_n${N}_overlap.json
timeout -s KILL 400 $TR --master-port 29903 bench.py --gpus $N --steps 8 --warmup 3 --impl reference > gpurun_out/bench_v2_n${N}_ref.json 2> gpurun_out/bench_v2_n${N}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_v2_n${N}_ref.json
