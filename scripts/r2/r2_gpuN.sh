#!/bin/bash
# multi-GPU check: cross-GPU tests, our arm and the R1 arm at N GPUs.  usage: r2_gpuN.sh N outdir [tests|notests] [r1|nor1]
N=${1:-2}; O=gpurun_out/${2:-r2n$N}
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ "${3:-tests}" = "tests" ]; then
timeout 900 python -m pytest tests/test_multigpu.py -q --maxfail=20 -p no:cacheprovider --timeout 240 > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multigpu.log
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_multigpu.log | tail -20
fi
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > $O/$2.json 2> $O/$2.err; echo "$2 rc=$?"; cut -c1-2000 $O/$2.json; }
run 29801 bench_ours --steps 20 --warmup 5
run 29802 bench_ours_200 --steps 200 --warmup 10
run 29803 bench_ours_nooverlap --steps 50 --warmup 5 --overlap 0 --skip-e2e
if [ "${4:-r1}" = "r1" ]; then run 29804 bench_r1 --impl nccl_cudnn --steps 20 --warmup 5; fi
run 29805 bench_fs --model fs --steps 300 --warmup 20
if [ "${4:-r1}" = "r1" ]; then run 29806 bench_fs_r1 --model fs --impl nccl_cudnn --steps 300 --warmup 20; fi
for f in $O/*.err; do echo "== $f"; grep -v "Warning\|warn" $f | tail -n 3 | cut -c1-300; done
