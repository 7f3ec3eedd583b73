set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_soup_lbvh.py tests/test_quirks.py -m gpu -x -q > gpurun_out/g10_new.log 2>&1; echo "new rc=$?" >> gpurun_out/g10_new.log
tail -15 gpurun_out/g10_new.log
timeout 900 python bench.py --workload soup_1M --steps 5 --warmup 3 > gpurun_out/g10_soup1m.json 2> gpurun_out/g10_soup1m.err; echo "soup1m rc=$?"
tail -c 800 gpurun_out/g10_soup1m.err; cut -c1-2500 gpurun_out/g10_soup1m.json
timeout 900 python bench.py --workload soup_1M --steps 5 --warmup 3 --soup-builder host --no-cpu-baseline > gpurun_out/g10_soup1m_host.json 2> gpurun_out/g10_soup1m_host.err; echo "soup1m host rc=$?"
cut -c1-1500 gpurun_out/g10_soup1m_host.json
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/g10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g10_tests.log
tail -8 gpurun_out/g10_tests.log
