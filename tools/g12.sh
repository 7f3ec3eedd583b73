set -x
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/g12_mem.txt
bash tools/prof.sh r02f traffic
head -40 gpurun_out/digest_r02f.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/g12_bench.json 2> gpurun_out/g12_bench.err; echo "bench rc=$?"
timeout 1200 python bench.py --workload soup_10M --steps 5 --warmup 3 --cpu-seconds 6 > gpurun_out/g12_soup10m.json 2> gpurun_out/g12_soup10m.err; echo "soup10m rc=$?"
tail -c 600 gpurun_out/g12_soup10m.err
bash tools/prof.sh r02f soup_10M
cat gpurun_out/digest_soup_10M_r02f.txt
timeout 1500 python bench.py --workload soup_100M --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/g12_soup100m.json 2> gpurun_out/g12_soup100m.err; echo "soup100m rc=$?"
tail -c 600 gpurun_out/g12_soup100m.err
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_r02f_final.json
