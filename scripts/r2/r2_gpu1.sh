#!/bin/bash
# round-2 single-GPU check: full GPU test suite, our arm, the R1 arm (cuDNN + NCCL + fused Adam), FreeSurfer MLP
O=gpurun_out/${1:-r2a}
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
timeout 600 python bench.py --steps 40 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"; cat $O/bench_ours.json
COINN_R1_LAUNCHES=1 timeout 600 python bench.py --impl nccl_cudnn --steps 40 --warmup 5 > $O/bench_r1.json 2> $O/bench_r1.err; echo "r1 rc=$?"; cat $O/bench_r1.json
timeout 300 python bench.py --model fs --steps 300 --warmup 20 > $O/bench_fs.json 2> $O/bench_fs.err; echo "fs rc=$?"; cat $O/bench_fs.json
timeout 300 python bench.py --model fs --impl nccl_cudnn --steps 300 --warmup 20 > $O/bench_fs_r1.json 2> $O/bench_fs_r1.err; echo "fs r1 rc=$?"; cat $O/bench_fs_r1.json
tail -5 $O/*.err
