for d in 16 8 5 3; do python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --max-depth $d 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('max_depth', $d, 'ms', round(d['ms_per_step'],2), 'rays/path', round(d['rays_per_path'],4), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, d['counts_rank0']['trace_launches'])"; done
