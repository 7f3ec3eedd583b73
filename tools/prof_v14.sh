# usage: prof_v14.sh shade|trace   (one capture per gpurun call: the merged gpurun_out/ is limited to 64 MiB)
B="python bench.py --steps 1 --warmup 0 --spp-per-step 1 --no-cpu-baseline"
$B > gpurun_out/plain_v14.json 2> gpurun_out/plain_v14.err || exit 1
if [ "$1" = shade ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_shade -c 18 -o gpurun_out/prof_shade_v14 -f $B > gpurun_out/ncu_s14.log 2>&1
else
  ncu --set full --clock-control none --import-source on -k regex:k_trace_fused -c 2 -o gpurun_out/prof_trace_v14 -f $B > gpurun_out/ncu_t14.log 2>&1
fi
ls -la gpurun_out/*v14*
