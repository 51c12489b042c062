set -x
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_ops_gpu.py -q -k "conv1_fused" > gpurun_out/pytest_c1f.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|assert " gpurun_out/pytest_c1f.log | head -30 | cut -c1-250
timeout -s KILL 240 python scripts/prof_c1f.py
