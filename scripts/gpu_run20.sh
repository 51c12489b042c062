set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "tcgen05_conv3d_matches or conv1 or native_vbm" > gpurun_out/pytest_2cta.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_2cta.log | head -5
python scripts/prof_conv.py time
COINN_HALO_CTAS=1 python scripts/prof_conv.py time
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --input-dtype fp32 --skip-e2e > gpurun_out/bench_v9.log 2>&1; tail -1 gpurun_out/bench_v9.log | cut -c1-250
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_v9_bf16in.log 2>&1; tail -1 gpurun_out/bench_v9_bf16in.log | cut -c1-250
for tool in memcheck racecheck; do
  timeout -s KILL 500 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_ops_gpu.py -q -x -k "count_binary or count_confusion or softmax or orthogonalize or fused_local_step or bn_relu_pool_block or tcgen05_gemm" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Error" gpurun_out/sanitize_$tool.log | tail -4
done
