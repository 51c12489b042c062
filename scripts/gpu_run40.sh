timeout -s KILL 200 python -m pytest tests/test_ops_gpu.py -q -k "conv1_fused or direct_grad or vbm" 2>&1 | tail -2
timeout -s KILL 240 python scripts/prof_c1f.py 2>&1 | head -4
