#!/bin/bash
# 2-GPU check of the shared-memory control plane: driver-config bench (20/5) with it and, as control, without it
O=gpurun_out/${1:-r2n2c}
mkdir -p $O
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" > $O/$2.json 2> $O/$2.err; echo "$2 rc=$?"; python - <<PY
import json
try:
    d=json.load(open('$O/$2.json')); e=d.get('e2e') or {}
    print('   ms/step', round(d['ms_per_step'],4), 'e2e', round(e.get('ms_per_step',0),4), 'samples/s', round(d['value']), round(e.get('value',0)))
except Exception as ex: print('   ERR', ex)
PY
}
run 29801 bench_shm --steps 20 --warmup 5
COINN_CTL_SHM=0 run 29802 bench_gloo --steps 20 --warmup 5
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref_n1.json 2> $O/ref_n1.err; echo "ref rc=$?"; cut -c1-500 $O/ref_n1.json; grep -v Warn $O/ref_n1.err | tail -n 2 | cut -c1-300
