mkdir -p gpurun_out
echo "default"; python scripts/prof_conv.py time
echo "FULLPIX=1"; COINN_HALO_FULLPIX=1 python scripts/prof_conv.py time
echo "bench FULLPIX=1"; COINN_HALO_FULLPIX=1 timeout -s KILL 600 python bench.py --skip-e2e 2>/dev/null | cut -c1-250
