set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "s19 or s3 or s17" > gpurun_out/g6_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/g6_parity.log
tail -3 gpurun_out/g6_parity.log
BENCH_ARGS="--no-extras" bash tools/run_ab.sh sync128 sync256 sync512 t256 > gpurun_out/g6_ab.log 2>&1
cat gpurun_out/g6_ab.log
