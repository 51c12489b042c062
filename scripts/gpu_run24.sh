set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -x -k "conv1_toeplitz" > gpurun_out/pytest_c1t.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/pytest_c1t.log | cut -c1-220
timeout -s KILL 300 python scripts/prof_bn.py fwdonly
