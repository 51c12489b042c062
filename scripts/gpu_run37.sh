mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "wgrad or conv3d or vbm or fused_bn" 2>&1 | tail -3 | cut -c1-200
echo "M64=1"; python scripts/prof_conv.py time
echo "M64=0"; COINN_WGRAD_M64=0 python scripts/prof_conv.py time
timeout -s KILL 240 python scripts/prof_c1f.py 2>&1 | sed -n 2,4p
timeout -s KILL 600 python bench.py --skip-e2e 2>/dev/null | cut -c1-250
