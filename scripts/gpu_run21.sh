set -x
mkdir -p gpurun_out
python scripts/prof_misc.py
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tn_kernel|fused_reduce_opt_kernel" -s 6 -c 2 -o gpurun_out/prof_gemm_adam -f python scripts/prof_misc.py > gpurun_out/ncu_misc.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv3d_halo_kernel|conv3d_wgrad_halo" -s 3 -c 3 -o gpurun_out/prof_halo2 -f python scripts/prof_conv.py prof > gpurun_out/ncu_halo2.log 2>&1; echo "ncu rc=$?"
