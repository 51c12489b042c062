set -x
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --master-port 29901 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_v3_n$N.json 2> gpurun_out/bench_v3_n$N.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench_v3_n$N.json
timeout -s KILL 300 $TR --master-port 29903 bench.py --gpus $N --steps 8 --warmup 3 --impl reference > gpurun_out/bench_v3_n${N}_ref.json 2> gpurun_out/bench_v3_n${N}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_v3_n${N}_ref.json
timeout -s KILL 200 $TR --master-port 29905 bench.py --gpus $N --model fs --steps 500 --warmup 20 --skip-e2e > gpurun_out/bench_v3_fs_n$N.json 2> gpurun_out/bench_v3_fs_n$N.err; cut -c1-300 gpurun_out/bench_v3_fs_n$N.json
