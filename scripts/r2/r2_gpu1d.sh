#!/bin/bash
# single-GPU: whole GPU suite (incl. new fp8 / lowrank / nativize / fused linear tests), FS + VBM benches
O=gpurun_out/${1:-r2d}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest.log | tail -40
timeout 300 python bench.py --steps 40 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"; cut -c1-600 $O/bench_ours.json
timeout 200 python bench.py --model fs --steps 300 --warmup 20 > $O/bench_fs.json 2> $O/bench_fs.err; echo "fs rc=$?"; cut -c1-400 $O/bench_fs.json
