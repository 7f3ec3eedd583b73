set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/g8_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/g8_parity.log
tail -5 gpurun_out/g8_parity.log
BENCH_ARGS="--no-extras" bash tools/run_ab.sh nosync128 sync256 > gpurun_out/g8_ab.log 2>&1
cat gpurun_out/g8_ab.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --opt env_nee_table=0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('env_nee_table=0', 'ms', round(d['ms_per_step'],2), d['stage_ms_per_step'])" > gpurun_out/g8_noenv.log 2>&1
cat gpurun_out/g8_noenv.log
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_shade -c 41 --csv --log-file gpurun_out/shade_launches_r02e.csv $BT > /dev/null 2>&1
