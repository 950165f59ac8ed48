#!/bin/bash
# ncu --set full captures (with source) of the dominant forest kernel launches of C3 (packed, self-looping leaves) and C1
# (resident fp32 forest); the small APPLY launch that measures path lengths is skipped by the kernel-name filter.
# usage (through gpurun): bash tools/ncu_main_kernels.sh <tag>
T=${1:-r02g}; O=gpurun_out
timeout 280 ncu --set full --clock-control none --import-source on -k regex:hm_forest_kernel --launch-skip 1 -c 1 -o $O/${T}_forest_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i $O/${T}_forest_c3.ncu-rep --page raw --csv > $O/${T}_forest_c3_ncu_raw.csv 2>/dev/null
timeout 280 ncu --set full --clock-control none --import-source on -k regex:hm_forest_kernel --launch-skip 1 -c 1 -o $O/${T}_forest_c1 -f python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i $O/${T}_forest_c1.ncu-rep --page raw --csv > $O/${T}_forest_c1_ncu_raw.csv 2>/dev/null
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
