for cfg in "0-15 5" "0-3 5" "0-3 3" "0-3 8" "0-1 5"; do
  set -- $cfg
  taskset -c $1 python bench.py --no-variants --no-cpu-baseline --steps 3 --e2e-inflight $2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cpus $1 inflight $2', 'value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],1), 'cpu_ms', round(d['e2e']['host_cpu_ms_per_request'],1))
"
done
