set -x
B="python bench.py --steps 1 --warmup 0 --spp-per-step 1 --no-cpu-baseline"
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 6 --warmup 3 > gpurun_out/bench14.json 2> gpurun_out/bench14.err
$B > gpurun_out/plain_v9.json 2> gpurun_out/plain_v9.err && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_v9.csv $B > gpurun_out/ncu_l9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_fused -c 2 -o gpurun_out/prof_trace_v9 -f $B > gpurun_out/ncu_t9.log 2>&1
ls -la gpurun_out/*v9*
