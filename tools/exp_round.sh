run() { echo "== $1"; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], l.get('parity',{}).get('ok'))"; }
V=/root/repo/variants
python -m pytest tests/test_packed_gpu.py tests/test_rf_gpu.py tests/test_api_gpu.py tests/test_selection_gpu.py -x -q -m gpu 2>&1 | tail -3
run c1 c1 A=1
run c1-noleafsmem c1 HM_B200_NO_LEAF_SMEM=1
run c1-base c1 HM_B200_LIB=$V/libhm_b200_base.so
run c1 c1 A=1
run c3 c3 A=1
