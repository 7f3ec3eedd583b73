set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hit_records or s19 or s3" > gpurun_out/g3_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/g3_parity.log
tail -3 gpurun_out/g3_parity.log
bash tools/run_ab.sh both r8 r16 r20 t4 t12 > gpurun_out/g3_ab.log 2>&1
cat gpurun_out/g3_ab.log
