run() { echo "== $1"; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], l.get('parity',{}).get('ok'))"; }
V=/root/repo/variants
run c1 c1 A=1
run c1-nokeep c1 HM_B200_LIB=$V/libhm_b200_nokeep.so
run c1-base c1 HM_B200_LIB=$V/libhm_b200_base.so
run c1 c1 A=1
run c1-nokeep c1 HM_B200_LIB=$V/libhm_b200_nokeep.so
