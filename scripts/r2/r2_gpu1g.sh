#!/bin/bash
O=gpurun_out/${1:-r2g}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout 300 -k "nativize or bucketed_overlap_with_direct" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|^E   " $O/pytest.log | cut -c1-1200 | tail -30
