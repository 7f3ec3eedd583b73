# Developer tool: A/B of the in-tree libtcpt.so against saved builds (toy_cpu_pathtracing_b200/lib/variants/libtcpt_<name>.so), interleaved.  usage: run_ab.sh name...
for rep in 1 2; do
for v in "" "$@"; do
  if [ -n "$v" ]; then export TCPT_LIB=$PWD/toy_cpu_pathtracing_b200/lib/variants/libtcpt_$v.so; else unset TCPT_LIB; fi
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant', '$v' or 'tree', 'ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
done
