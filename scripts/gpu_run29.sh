mkdir -p gpurun_out
for d in 2 3 5; do echo "dbg=$d"; COINN_C1F_DEBUG=$d timeout -s KILL 120 python scripts/prof_c1f_dbg.py; done
