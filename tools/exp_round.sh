run() { echo "== $1"; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], (l.get('e2e') or {}).get('ms_per_step'), l.get('parity',{}).get('ok'))"; }
V=/root/repo/variants
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_gpu_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r02m_gpu_tests.log)"
HM_B200_LIB=hypermapper_b200/csrc/libhm_b200_checked.so python tools/checked_run.py > gpurun_out/r02m_checked_run.log 2>&1; echo "checked rc=$? $(tail -1 gpurun_out/r02m_checked_run.log)"
run c3 c3 A=1
run c5rf c5rf A=1
run c3f32 c3f32 A=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02m_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -c hm_forest_scan gpurun_out/r02m_launches_c3.csv
