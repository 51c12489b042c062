#!/bin/bash
# the one 8-GPU run of the round: e2e scaling of the headline, R1 at N=8, cross-GPU tests, reduce sweep
N=8; O=gpurun_out/${1:-r2n8}
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > $O/$2.json 2> $O/$2.err; echo "$2 rc=$?"; cut -c1-1800 $O/$2.json; }
run 29801 bench_ours --steps 20 --warmup 5
run 29802 bench_ours_200 --steps 200 --warmup 10
run 29804 bench_r1 --impl nccl_cudnn --steps 20 --warmup 5
run 29805 bench_fs --model fs --steps 300 --warmup 20
run 29807 bench_ours_again --steps 20 --warmup 5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29810 bench_reduce.py --max-mb 64 --iters 4 --burst 5 > $O/reduce_sweep.log 2> $O/reduce_sweep.err; echo "sweep rc=$?"; tail -n 70 $O/reduce_sweep.log | cut -c1-260
cp gpurun_out/reduce_sweep_n8.json $O/ 2>/dev/null
timeout 420 python -m pytest tests/test_multigpu.py -q --maxfail=20 -p no:cacheprovider --timeout 200 -k "fused_reduce_all_variants or protocol_over_nvlink or rankdad or powersgd_after or watchdog or overlap_in_graph or sixteen_bit" > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_multigpu.log | tail -20
for f in $O/*.err; do echo "== $f"; grep -v "Warning\|warn\|symm\." $f | tail -n 3 | cut -c1-300; done
