#!/bin/bash
# end-of-round evidence on one B200: full GPU test suite, checked build, every bench workload, launch lists, ncu captures
# usage (through gpurun): bash tools/final_1gpu.sh <tag>      outputs under gpurun_out/
T=${1:-r02f}; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
HM_B200_LIB=hypermapper_b200/csrc/libhm_b200_checked.so python tools/checked_run.py > $O/${T}_checked_run.log 2>&1; echo "checked rc=$? $(tail -1 $O/${T}_checked_run.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/${T}_smoke.log)"
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
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_c3_reference_arm.json 2> $O/${T}_bench_c3_reference_arm.err; tail -c 400 $O/${T}_bench_c3_reference_arm.json
python tools/pareto_timeline.py > $O/${T}_pareto_timeline.txt 2>&1
# launch lists (same commands, under ncu: numbers printed by these runs are not bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c1.csv python bench.py --workload c1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:hm_forest_kernel --launch-skip 1 -c 1 -o $O/${T}_forest_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i $O/${T}_forest_c3.ncu-rep --page raw --csv > $O/${T}_forest_c3_ncu_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:hm_forest_kernel --launch-skip 1 -c 1 -o $O/${T}_forest_c1 -f python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i $O/${T}_forest_c1.ncu-rep --page raw --csv > $O/${T}_forest_c1_ncu_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:hm_px_stream -c 1 -o $O/${T}_pareto_stream -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i $O/${T}_pareto_stream.ncu-rep --page raw --csv > $O/${T}_pareto_stream_ncu_raw.csv 2>/dev/null
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
