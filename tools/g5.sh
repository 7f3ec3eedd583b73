set -x
mkdir -p gpurun_out
BT="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 1 --launch-count 1 -o gpurun_out/prof_shade1_16spp_r02c -f $BT > gpurun_out/ncu_s1_r02c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 5 --launch-count 1 -o gpurun_out/prof_shade5_16spp_r02c -f $BT > gpurun_out/ncu_s5_r02c.log 2>&1
ls -la gpurun_out/*r02c*
