# Developer tool (run under gpurun): ncu --set full of the miss-bucket shading launches (bounce 0 and 1) of one default 16-spp step,
# digested on the box (the reports are too large to travel).  usage: gprof.sh <tag>
tag="${1:-x}"
mkdir -p gpurun_out
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
$BT > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_shade<7" -c 2 -o /tmp/prof_miss -f $BT > gpurun_out/ncu_miss_$tag.log 2>&1
if [ ! -f /tmp/prof_miss.ncu-rep ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 7 --launch-count 1 -o /tmp/prof_miss -f $BT >> gpurun_out/ncu_miss_$tag.log 2>&1
fi
ls -la /tmp/prof_miss.ncu-rep
python tools/ncu_summary.py /tmp/prof_miss.ncu-rep > gpurun_out/miss_summary_$tag.txt 2>&1
for i in 0 1; do python tools/ncu_by_source.py /tmp/prof_miss.ncu-rep toy_cpu_pathtracing_b200/lib/libtcpt.so $i 60 >> gpurun_out/miss_by_source_$tag.txt 2>&1; done
ncu -i /tmp/prof_miss.ncu-rep --page details > gpurun_out/miss_details_$tag.txt 2>&1
