set -x
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused_kernel" -c 3 -o gpurun_out/prof_c1f -f python scripts/prof_c1f.py prof > gpurun_out/ncu_c1f.log 2>&1; echo "ncu rc=$?"
