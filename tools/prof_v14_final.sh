# launch list of the default bench command (shortened to 2 steps), then a full capture of the dominant kernel (final state of round 1)
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/plain_v14.json 2> gpurun_out/plain_v14.err || exit 1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v14.csv $B > gpurun_out/ncu_l13.log 2>&1
B1="python bench.py --steps 1 --warmup 0 --spp-per-step 1 --no-cpu-baseline"
$B1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace_fused -c 2 -o gpurun_out/prof_trace_v14 -f $B1 > gpurun_out/ncu_t13.log 2>&1
ls -la gpurun_out/*v14*
