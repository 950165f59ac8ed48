#!/bin/bash
# every bench workload once on one B200 (parity gate on), JSON lines under gpurun_out/<tag>_bench_<workload>.json
T=${1:-r02n}; O=gpurun_out
line() { python -c "
import json,sys
try:
    l=json.loads(open('$1').read().strip().splitlines()[-1]); print('%s value=%.4g ms=%.4f kernel_ms=%.4f e2e=%.4g parity=%s' % ('$1', l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['e2e']['value'] if l.get('e2e') else 0, (l.get('parity') or {}).get('ok')))
except Exception as e: print('$1 no line', e)
"; }
python bench.py > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err; line $O/${T}_bench_c3.json
for wl in c1 c2 c4 c5rf c5gp c3f32 c3raw c3pool; do python bench.py --workload $wl > $O/${T}_bench_$wl.json 2> $O/${T}_bench_$wl.err; line $O/${T}_bench_$wl.json; done
python bench.py --workload c1 --fused-topk --no-e2e > $O/${T}_bench_c1_fused.json 2> $O/${T}_bench_c1_fused.err; line $O/${T}_bench_c1_fused.json
python bench.py --fused-topk --no-e2e > $O/${T}_bench_c3_fused.json 2> $O/${T}_bench_c3_fused.err; line $O/${T}_bench_c3_fused.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_c3_reference_arm.json 2> $O/${T}_bench_c3_reference_arm.err; tail -c 300 $O/${T}_bench_c3_reference_arm.json
