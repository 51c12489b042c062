#!/bin/bash
O=gpurun_out/${1:-r2b}
mkdir -p $O
timeout 600 python scripts/diag_determinism.py > $O/diag.log 2>&1; echo "diag rc=$?"; cat $O/diag.log | tail -120
timeout 600 python -m pytest tests/test_hardening_gpu.py -q -k "converges" -p no:cacheprovider > $O/pytest_conv.log 2>&1; tail -15 $O/pytest_conv.log
COINN_R1_LAUNCHES=1 timeout 600 python bench.py --impl nccl_cudnn --steps 40 --warmup 5 > $O/bench_r1.json 2> $O/bench_r1.err; echo "r1 rc=$?"; cat $O/bench_r1.json; tail -n 5 $O/bench_r1.err
