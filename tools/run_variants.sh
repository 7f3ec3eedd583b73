for v in "" pfl1; do
  if [ -n "$v" ]; then export TCPT_LIB=$PWD/toy_cpu_pathtracing_b200/lib/variants/libtcpt_$v.so; else unset TCPT_LIB; fi
  python bench.py --steps 4 --warmup 3 --spp-per-step 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=8; print('variant', '$v' or 'base', 'ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']), {k: round(v/s,3) for k,v in d['stage_ms_per_step'].items()})"
done
