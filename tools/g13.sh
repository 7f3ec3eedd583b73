set -x
mkdir -p gpurun_out
BENCH_ARGS="--no-extras" bash tools/run_ab.sh lam640 lam768 > gpurun_out/g13_ab.log 2>&1
cat gpurun_out/g13_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_quirks.py tests/test_soup_lbvh.py -m gpu -x -q > gpurun_out/g13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g13_tests.log
tail -5 gpurun_out/g13_tests.log
