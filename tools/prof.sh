# Profiling recipe of profiles/ (run under gpurun; one capture kind per call: the merged gpurun_out/ is limited to 64 MiB).
# usage: prof.sh <tag> launches|trace|shade|traffic|soup_1M|soup_10M|soup_100M
#   launches : ncu launch list (duration, active lanes, DRAM bytes per launch) of the default bench command shortened to 2 steps,
#              after the same command has run once WITHOUT ncu (its JSON line is the twin the kernel shares are checked against)
#   trace    : ncu --set full of two k_trace_fused launches (bounce 0 and 1) of a 1-spp pass
#   shade    : ncu --set full of the first 18 k_shade launches (all buckets of bounce 0 and 1) of a 1-spp pass
# digest afterwards with tools/ncu_summary.py / tools/ncu_by_source.py and copy what should be judged into profiles/
tag="$1"; kind="$2"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
B1="python bench.py --steps 1 --warmup 0 --spp-per-step 1 --no-cpu-baseline"
case "$kind" in
  launches)
    $B > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.err || exit 1
    ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l_$tag.log 2>&1 ;;
  trace)
    $B1 > /dev/null 2>&1 || exit 1
    ncu --set full --clock-control none --import-source on -k regex:k_trace_fused -c 2 -o gpurun_out/prof_trace_$tag -f $B1 > gpurun_out/ncu_t_$tag.log 2>&1 ;;
  shade)
    $B1 > /dev/null 2>&1 || exit 1
    ncu --set full --clock-control none --import-source on -k regex:k_shade -c 18 -o gpurun_out/prof_shade_$tag -f $B1 > gpurun_out/ncu_s_$tag.log 2>&1 ;;
  traffic)   # per-launch DRAM bytes, time, FP32 op counts and lanes of ONE default step -> tools/ncu_digest.py -> profiles/ncu_traffic.json
    BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
    $BT > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.err || exit 1
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum --clock-control none -k regex:k_ -c 260 --csv --log-file gpurun_out/traffic_$tag.csv $BT > gpurun_out/ncu_tr_$tag.log 2>&1
    python tools/ncu_digest.py gpurun_out/traffic_$tag.csv gpurun_out/plain_$tag.json $tag > gpurun_out/digest_$tag.txt 2>&1
    cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_$tag.json; cp profiles/${tag}_step_launches.json gpurun_out/ ;;
  soup_*)    # DRAM bytes of the traversal launches of one soup workload -> profiles/ncu_traffic.json["soups"]
    BS="python bench.py --workload $kind --steps 1 --warmup 0 --no-cpu-baseline"
    $BS > gpurun_out/plain_${kind}_$tag.json 2> gpurun_out/plain_${kind}_$tag.err || exit 1
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_trace_rays -c 8 --csv --log-file gpurun_out/traffic_${kind}_$tag.csv $BS > gpurun_out/ncu_${kind}_$tag.log 2>&1
    python tools/ncu_digest.py --soup gpurun_out/traffic_${kind}_$tag.csv gpurun_out/plain_${kind}_$tag.json $tag $kind > gpurun_out/digest_${kind}_$tag.txt 2>&1
    cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_$tag.json ;;
  *) echo "usage: prof.sh <tag> launches|trace|shade"; exit 2 ;;
esac
ls -la gpurun_out/*$tag*
