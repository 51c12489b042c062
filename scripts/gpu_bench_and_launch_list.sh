mkdir -p gpurun_out
timeout -s KILL 600 python bench.py > gpurun_out/bench_r42.json 2> gpurun_out/bench_r42.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r42.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 260 --csv --log-file gpurun_out/launches_v15.csv python bench.py --steps 4 --warmup 3 --graph 0 --skip-e2e > gpurun_out/ncu_v15.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_v15.csv 60 > gpurun_out/launches_v15.txt; head -42 gpurun_out/launches_v15.txt | cut -c1-140
