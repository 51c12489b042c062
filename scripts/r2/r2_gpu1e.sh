#!/bin/bash
O=gpurun_out/${1:-r2e}
mkdir -p $O
timeout 120 python scripts/debug_lbr.py > $O/debug_lbr.log 2>&1; cat $O/debug_lbr.log | tail -20
timeout 600 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout 300 -k "nativize or bucketed_overlap_with_direct or fused_linear_bn1d or powersgd_kernels or lowrank_factor or fp8 or native_fsnet" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest.log | tail -40
