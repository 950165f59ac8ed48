run() { echo "== $1"; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --no-e2e --steps 3 --warmup 2 --cpu-sample 50000 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['kernel_ms'], (l.get('parity') or {}).get('ok'))"; }
run c3-dfs c3 HM_B200_NODE_ORDER=dfs
run c3-bfs c3 A=1
run c5rf-dfs c5rf HM_B200_NODE_ORDER=dfs
run c5rf-bfs c5rf A=1
