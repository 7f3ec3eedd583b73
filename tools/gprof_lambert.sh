# Developer tool (run under gpurun): ncu --set full of the bounce-0 Lambert shading launch of one default 16-spp step, digested on the box.
tag="${1:-x}"
mkdir -p gpurun_out
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
$BT > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 5 --launch-count 1 -o /tmp/prof_lam -f $BT > gpurun_out/ncu_lam_$tag.log 2>&1
ls -la /tmp/prof_lam.ncu-rep
python tools/ncu_summary.py /tmp/prof_lam.ncu-rep > gpurun_out/lambert_summary_$tag.txt 2>&1
python tools/ncu_by_source.py /tmp/prof_lam.ncu-rep toy_cpu_pathtracing_b200/lib/libtcpt.so 0 90 >> gpurun_out/lambert_by_source_$tag.txt 2>&1
