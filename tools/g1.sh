set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/g1_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_intersection_ground_truth.py -m gpu -x -q > gpurun_out/g1_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/g1_parity.log
tail -5 gpurun_out/g1_parity.log
BENCH_ARGS="" bash tools/run_ab.sh b8 st16 nosort b6 > gpurun_out/g1_ab.log 2>&1
cat gpurun_out/g1_ab.log
timeout 1200 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/g1_full.log 2>&1; echo "full rc=$?" >> gpurun_out/g1_full.log
tail -15 gpurun_out/g1_full.log
