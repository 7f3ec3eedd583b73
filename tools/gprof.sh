set -x
mkdir -p gpurun_out
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 20 -o gpurun_out/prof_shade_all20_r02j -f $BT > gpurun_out/ncu_s20_r02j.log 2>&1
ls -la gpurun_out/prof_shade_all20_r02j.ncu-rep
