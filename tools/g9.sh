set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g9_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/g9_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/g9_multi.log
tail -15 gpurun_out/g9_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/g9_bench_n2.json 2> gpurun_out/g9_bench_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/g9_bench_n2.err
cat gpurun_out/g9_bench_n2.json | cut -c1-3000
