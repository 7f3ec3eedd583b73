set -x
mkdir -p gpurun_out
export TCPT_LIB=$PWD/toy_cpu_pathtracing_b200/lib/variants/libtcpt_sync512.so
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 1 --launch-count 1 -o gpurun_out/prof_shade1_sync512_r02d -f $BT > gpurun_out/ncu_s1_r02d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 5 --launch-count 1 -o gpurun_out/prof_shade5_sync512_r02d -f $BT > gpurun_out/ncu_s5_r02d.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_shade -c 41 --csv --log-file gpurun_out/shade_launches_sync512_r02d.csv $BT > /dev/null 2>&1
ls -la gpurun_out/*r02d*
