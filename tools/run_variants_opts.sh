# Developer tool: the default bench step under different libtcpt options (same binary).  usage: run_variants_opts.sh "name:opt=val,opt=val" ...
for spec in "$@"; do
  name="${spec%%:*}"; opts="${spec#*:}"; args=""
  if [ "$opts" != "$spec" ] && [ -n "$opts" ]; then for o in ${opts//,/ }; do args="$args --opt $o"; done; fi
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant', '$name', 'ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
