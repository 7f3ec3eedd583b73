# Developer tool (run under gpurun): ncu --set full of the first two k_trace_fused launches (camera rays; bounce-0 shadow rays + bounce-1
# extension rays) of one default 16-spp step, digested on the box.  usage: gprof_trace.sh <tag>
tag="${1:-x}"
mkdir -p gpurun_out
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
$BT > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace_fused -c 2 -o /tmp/prof_trace -f $BT > gpurun_out/ncu_trace_$tag.log 2>&1
ls -la /tmp/prof_trace.ncu-rep
python tools/ncu_summary.py /tmp/prof_trace.ncu-rep > gpurun_out/trace_summary_$tag.txt 2>&1
for i in 0 1; do python tools/ncu_by_source.py /tmp/prof_trace.ncu-rep toy_cpu_pathtracing_b200/lib/libtcpt.so $i 70 >> gpurun_out/trace_by_source_$tag.txt 2>&1; done
