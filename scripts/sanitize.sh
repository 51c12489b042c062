#!/usr/bin/env bash
# Race / memory checks of the hand-written kernels.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# single-GPU kernels under compute-sanitizer memcheck / racecheck / synccheck; the cross-GPU flag protocol is
# stress-tested by tests/test_multigpu.py (random per-rank delays, replica bit-equality).
set -x
mkdir -p gpurun_out
export COINN_CONV_IMPL=auto
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --error-exitcode 1 \
    python -m pytest tests/test_ops_gpu.py -q -x -k "count or softmax or orthogonalize or fused_local or bn_relu_pool_block or conv1_fwd" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
