set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g2_tests.log
tail -30 gpurun_out/g2_tests.log
bash tools/prof.sh r02a trace
