#!/bin/bash
O=gpurun_out/${1:-r2c}
mkdir -p $O
timeout 600 python scripts/diag_stats.py > $O/diag_stats.log 2>&1; echo "diag rc=$?"; cat $O/diag_stats.log | tail -60
