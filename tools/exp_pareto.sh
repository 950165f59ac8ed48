python -m pytest tests/test_pareto_gpu.py -x -q -m gpu 2>&1 | tail -3
HM_B200_LIB=hypermapper_b200/csrc/libhm_b200_checked.so python -m pytest tests/test_pareto_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/pareto_timeline.py > gpurun_out/r02_pareto_timeline_v6.txt 2>&1; tail -16 gpurun_out/r02_pareto_timeline_v5.txt
python bench.py --workload c4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], l['e2e']['ms_per_step'])"
python -m pytest tests/test_space_gpu.py -x -q -m gpu 2>&1 | tail -2
python -m pytest tests/test_selection_gpu.py -x -q -m gpu -k fused 2>&1 | tail -5
