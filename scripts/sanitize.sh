#!/usr/bin/env bash
# Race / memory checks of the hand-written kernels.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# 1. CUDA-core kernels under memcheck / racecheck / synccheck;
# 2. the tcgen05 / TMA / TMEM kernels (GEMM, halo + TMA-box convs, fused first block, per-tap wgrad, MX-FP8 GEMM and conv,
#    fused Linear+BN1d, low-rank kernels) under memcheck on small shapes (racecheck does not model the async proxy);
# 3. the cross-GPU kernels (fused reduce+Adam all variants, symmetric all-gather) under memcheck when >= 2 GPUs are visible.
# The flag protocol itself is stress-tested by tests/test_multigpu.py (random per-rank delays, replica bit-equality).
O=gpurun_out/sanitize
mkdir -p $O
export COINN_CONV_IMPL=auto
SAN="compute-sanitizer --error-exitcode 1 --launch-timeout 0"
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 600 $SAN --tool $tool \
    python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "count or softmax or orthogonalize or fused_local or bn_relu_pool_block or small_linear or fused_linear_bn1d or dad_reconstruct" \
    > $O/cudacore_$tool.log 2>&1
  echo "cuda-core $tool rc=$?"; tail -2 $O/cudacore_$tool.log
done
# tensor-core / TMA kernels: one small case of each family (memcheck replays every launch: keep it short)
timeout -s KILL 900 $SAN --tool memcheck \
  python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider \
  -k "(tcgen05_gemm and 128-128-64) or (tcgen05_conv3d_matches and (halo or tma) and shape0) or (conv3d_wgrad and halo and shape0) or (conv1_fused_block and shape0) or (wgrad_tap and shape0) or (mxfp8_block_scaled and 128-128-128) or (mxfp8_conv3d and shape0) or (powersgd_kernels and 2) or (lowrank_factor and 32-33)" \
  > $O/tensorcore_memcheck.log 2>&1
echo "tensor-core memcheck rc=$?"; tail -3 $O/tensorcore_memcheck.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout -s KILL 600 $SAN --tool memcheck --target-processes all \
    python -m pytest tests/test_multigpu.py -q -x -p no:cacheprovider -k "fused_reduce_all_variants or symm_allreduce" \
    > $O/crossgpu_memcheck.log 2>&1
  echo "cross-GPU memcheck rc=$?"; tail -3 $O/crossgpu_memcheck.log
fi
