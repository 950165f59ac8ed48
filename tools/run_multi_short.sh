#!/bin/bash
# end-of-round multi-GPU lines (the kernels that changed late in the round): tools/run_multi_short.sh <n_gpus> [outdir]
N=$1; OUT=${2:-gpurun_out}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > $OUT/r02_final_bench_${name}_${N}gpu.json 2> $OUT/r02_final_bench_${name}_${N}gpu.err; echo "$name rc=$? $(python -c "
import json,sys
try:
    l=json.loads(open('$OUT/r02_final_bench_${name}_${N}gpu.json').read().strip().splitlines()[-1]); print('%.4g %s ms=%.3f e2e=%.4g parity=%s' % (l['value'], l['scaling'], l['ms_per_step'], l['e2e']['value'], (l.get('parity') or {}).get('ok')))
except Exception as e: print('no line', e)
")"; }
run c3 --steps 5 --warmup 3
run c5rf_64M --workload c5rf --total-candidates 67108864 --steps 3 --warmup 3
run c4 --workload c4 --steps 10 --warmup 3
