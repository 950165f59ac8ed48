#!/bin/bash
# multi-GPU bench lines of one box: tools/run_multi.sh <n_gpus> <outdir>
N=$1; OUT=${2:-gpurun_out}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > $OUT/r02_bench_${name}_${N}gpu.json 2> $OUT/r02_bench_${name}_${N}gpu.err; echo "$name rc=$? $(python -c "
import json,sys
try:
    l=json.loads(open('$OUT/r02_bench_${name}_${N}gpu.json').read().strip().splitlines()[-1]); print('%.4g %s ms=%.3f e2e=%.4g' % (l['value'], l['scaling'], l['ms_per_step'], l['e2e']['value']))
except Exception as e: print('no line', e)
")"; }
run c3 --steps 5 --warmup 3
run c5rf_64M --workload c5rf --total-candidates 67108864 --steps 3 --warmup 3
run c5gp_64M --workload c5gp --total-candidates 67108864 --steps 3 --warmup 3
run c4 --workload c4 --steps 10 --warmup 3
NCCL_DEBUG=INFO $TR --master-port 29999 bench.py --gpus $N --candidates 1048576 --steps 1 --warmup 3 --no-e2e 2>&1 | grep -E "NCCL INFO (comm|Connected|ncclCommInitRank)" | head -4 > $OUT/r02_nccl_${N}gpu.txt
