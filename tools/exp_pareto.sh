python -m pytest tests/test_pareto_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/pareto_timeline.py > gpurun_out/r02_pareto_timeline_v3.txt 2>&1; tail -16 gpurun_out/r02_pareto_timeline_v3.txt
python bench.py --workload c4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], l['e2e']['ms_per_step'])"
HM_B200_LIB=hypermapper_b200/csrc/libhm_b200_checked.so python -m pytest tests/test_pareto_gpu.py -x -q -m gpu 2>&1 | tail -3
HM_B200_LIB=hypermapper_b200/csrc/libhm_b200_checked.so python tools/checked_run.py 2>&1 | tail -2
python -m pytest tests/test_packed_gpu.py tests/test_rf_gpu.py tests/test_api_gpu.py tests/test_gp_gpu.py tests/test_selection_gpu.py tests/test_space_gpu.py -x -q -m gpu 2>&1 | tail -3
for wl in c1 c2 c3; do python bench.py --workload $wl --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', l['ms_per_step'], l['roofline']['kernel_ms'], l.get('parity'))"; done
for p in 1024 128 64 32 16; do echo "piv1=$p"; HM_B200_PARETO_PIV1=$p python tools/pareto_timeline.py 2>&1 | grep -E "hm_px_stream|hm_px_prep|span"; done
