run() { echo "== $1"; shift; env "$@" python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'])"; }
run default A=1
run b512 HM_B200_LIB=/root/repo/variants/libhm_b200_b512.so
run b256 HM_B200_LIB=/root/repo/variants/libhm_b200_b256.so
HM_B200_LIB=/root/repo/variants/libhm_b200_b512.so python -m pytest tests/test_packed_gpu.py -x -q -m gpu 2>&1 | tail -2
