set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_multigpu.py -q > gpurun_out/pytest_multigpu.log 2>&1; echo "multigpu rc=$?"
grep -E "passed|failed|Error|assert|FAILED|KeyError" gpurun_out/pytest_multigpu.log | head -20
