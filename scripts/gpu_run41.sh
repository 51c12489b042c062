timeout -s KILL 200 python -m pytest tests/test_ops_gpu.py -q -k "wgrad_tap" 2>&1 | tail -8 | cut -c1-200
timeout -s KILL 120 python scripts/prof_wgrad45.py
