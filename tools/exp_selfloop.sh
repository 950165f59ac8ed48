run() { echo "== $1"; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], l.get('parity',{}).get('ok'))"; }
python -m pytest tests/test_packed_gpu.py tests/test_rf_gpu.py -x -q -m gpu 2>&1 | tail -3
run wl2 c3 A=1
run wl1 c3 HM_B200_LIB=/root/repo/variants/libhm_b200_wl1.so
run wl3 c3 HM_B200_LIB=/root/repo/variants/libhm_b200_wl3.so
run c5rf-wl2 c5rf A=1
run c5rf-wl1 c5rf HM_B200_LIB=/root/repo/variants/libhm_b200_wl1.so
run c3f32-wl2 c3f32 A=1
run c1-wl2 c1 A=1
run c1-wl1 c1 HM_B200_LIB=/root/repo/variants/libhm_b200_wl1.so
