set -x
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 450 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 4 --warmup 3 --graph 0 --skip-e2e > gpurun_out/ncu_v10.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_v10.csv 45 | tee gpurun_out/launches_v10.txt
