set -x
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm or linear or checkpoint" > gpurun_out/pytest_gemm.log 2>&1; echo "gemm rc=$?"
tail -25 gpurun_out/pytest_gemm.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_torchmodules.csv python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_bench.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_torchmodules.csv')))
hdr = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr+1:]:
    try: agg[r[kn][:90]] += float(r[mv].replace(',','')); cnt[r[kn][:90]] += 1
    except Exception: pass
tot = sum(agg.values())
print('total ns', tot)
for k,v in agg.most_common(25): print(f'{v/tot*100:6.2f}% {v/1e3:10.1f}us x{cnt[k]:4d} {k}')
PY
