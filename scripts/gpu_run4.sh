set -x
mkdir -p gpurun_out
python scripts/debug_linear.py 2>&1 | tail -8
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py -q -k "conv1 or bn_relu or native_vbm or tcgen05_conv3d" > gpurun_out/pytest_conv.log 2>&1; echo "conv rc=$?"
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_conv.log | head -40
