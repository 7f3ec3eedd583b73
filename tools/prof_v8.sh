set -x
B="python bench.py --steps 1 --warmup 0 --spp-per-step 1 --no-cpu-baseline"
$B > gpurun_out/plain_v8.json 2> gpurun_out/plain_v8.err && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/launches_v8.csv $B > gpurun_out/ncu_l8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 16 -o gpurun_out/prof_shade_v8 -f $B > gpurun_out/ncu_s8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -c 2 -o gpurun_out/prof_trace_v8 -f $B > gpurun_out/ncu_t8.log 2>&1
ls -la gpurun_out/*v8*
