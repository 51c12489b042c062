mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 330 --csv --log-file gpurun_out/launches_v14.csv python bench.py --steps 4 --warmup 3 --graph 0 --skip-e2e > gpurun_out/ncu_v14.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_v14.csv 60 > gpurun_out/launches_v14.txt; head -45 gpurun_out/launches_v14.txt | cut -c1-140
