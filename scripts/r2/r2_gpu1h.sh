#!/bin/bash
# final single-GPU validation: whole GPU suite, driver-config bench (20/5), steady-state bench, config-4 bench, ncu of the new kernels
O=gpurun_out/${1:-r2h}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest.log | tail -20
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench_ours_20.json 2> $O/bench_ours_20.err; echo "ours20 rc=$?"; cut -c1-400 $O/bench_ours_20.json
timeout 200 python bench.py --steps 200 --warmup 10 --skip-e2e > $O/bench_ours_200.json 2> $O/bench_ours_200.err; echo "ours200 rc=$?"; cut -c1-300 $O/bench_ours_200.json
timeout 200 python bench.py --dtype fp8 --steps 40 --warmup 5 --skip-e2e > $O/bench_fp8.json 2> $O/bench_fp8.err; echo "fp8 rc=$?"; cut -c1-300 $O/bench_fp8.json; tail -n 3 $O/bench_fp8.err
for k in conv_fp8:conv3d_mxfp8_kernel conv_fp8:conv3d_halo gemm_fp8:gemm_mxfp8_kernel lbr:linear_small_fwd_kernel; do
  w=${k%%:*}; n=${k##*:}
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$n -c 1 -f -o $O/ncu_$n python scripts/ncu_r2_driver.py $w > $O/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
done
ls -la $O | tail -12
