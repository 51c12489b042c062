set -x
mkdir -p gpurun_out
python scripts/debug_graph.py 2>&1 | tail -6
python scripts/prof_conv.py time
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo_kernel -s 2 -c 2 -o gpurun_out/prof_halo -f python scripts/prof_conv.py prof > gpurun_out/ncu_halo.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_wgrad_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad -f python scripts/prof_conv.py prof > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
