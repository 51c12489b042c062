#!/bin/bash
# bounded sanitizer pass (GPU budget): memcheck over one small case of every tensor-core / TMA kernel family + the CUDA-core kernels
O=gpurun_out/sanitize
mkdir -p $O
SAN="compute-sanitizer --error-exitcode 1 --launch-timeout 0"
timeout -s KILL 170 $SAN --tool memcheck python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider \
  -k "(tcgen05_gemm and 128-128-64) or (mxfp8_block_scaled and 128-128-128) or (mxfp8_conv3d and shape0) or (powersgd_kernels and 2) or (lowrank_factor and 32-33) or dad_reconstruct or (fused_linear_bn1d and 2-9-5)" \
  > $O/tensorcore_memcheck.log 2>&1; echo "tensor-core memcheck rc=$?"; tail -4 $O/tensorcore_memcheck.log | cut -c1-300
timeout -s KILL 120 $SAN --tool memcheck python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider \
  -k "(tcgen05_conv3d_matches and halo2 and 16-32-shape0) or (conv3d_wgrad_matches and halo and 16-32-shape0) or (conv1_fused_block and shape0)" \
  > $O/conv_memcheck.log 2>&1; echo "conv memcheck rc=$?"; tail -4 $O/conv_memcheck.log | cut -c1-300
timeout -s KILL 60 $SAN --tool racecheck python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "small_linear or (fused_linear_bn1d and 8-64-32) or softmax" \
  > $O/cudacore_racecheck.log 2>&1; echo "cuda-core racecheck rc=$?"; tail -3 $O/cudacore_racecheck.log | cut -c1-300
