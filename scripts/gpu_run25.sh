set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -x -k "conv1_toeplitz" > gpurun_out/pytest_c1t.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/pytest_c1t.log | cut -c1-220
timeout -s KILL 300 python scripts/prof_bn.py fwdonly
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_toeplitz" -c 1 -o gpurun_out/prof_c1t -f python scripts/prof_bn.py prof_t > gpurun_out/ncu_c1t.log 2>&1; echo "ncu rc=$?"
