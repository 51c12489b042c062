set -x
nvidia-smi -L | head -3
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" 
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench1.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench1_ref.log 2>&1; echo "ref rc=$?"; tail -3 gpurun_out/bench1_ref.log
cat gpurun_out/smoke.log | tail -5
