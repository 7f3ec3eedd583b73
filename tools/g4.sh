set -x
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 > gpurun_out/g4_bench.json 2> gpurun_out/g4_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/g4_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g4_ref.json 2> gpurun_out/g4_ref.err; echo "ref rc=$?"
bash tools/prof.sh r02b traffic
cat gpurun_out/digest_r02b.txt | head -60
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/g4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g4_tests.log
tail -8 gpurun_out/g4_tests.log
