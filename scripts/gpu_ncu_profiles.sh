set -x
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused_kernel" -s 2 -c 1 -o gpurun_out/prof_c1f_bwd -f python scripts/prof_conv.py prof > gpurun_out/ncu_c1f_bwd.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv3d_halo_kernel|conv3d_wgrad_halo" -s 3 -c 3 -o gpurun_out/prof_halo3 -f python scripts/prof_conv.py prof > gpurun_out/ncu_halo3.log 2>&1; echo "ncu rc=$?"
