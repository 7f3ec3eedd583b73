for k in 1 2 5 20; do python bench.py --steps $k --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('steps', $k, 'ms', round(d['ms_per_step'],2), 'sum_stages', round(sum(s.values()),2), 'Mrays/s', round(d['value']))"; done
