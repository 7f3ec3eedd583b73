set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_soup_lbvh.py -m gpu -x -q > gpurun_out/g11_soup.log 2>&1; echo "soup rc=$?" >> gpurun_out/g11_soup.log
tail -12 gpurun_out/g11_soup.log
for leaf in 1 2 4; do
timeout 600 python bench.py --workload soup_1M --steps 5 --warmup 3 --no-cpu-baseline --opt soup_leaf=$leaf 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('leaf $leaf', 'inc', round(d['value']), 'coh', round(d['coherent_closest']['mrays_per_s']), 'any', round(d['incoherent_anyhit']['mrays_per_s']), 'B/T', round(d['incoherent_closest']['box_tests_per_ray']), round(d['incoherent_closest']['tri_tests_per_ray']), d['device_build'], 'depth', d['bvh_depth'])"
done > gpurun_out/g11_leaf.log 2>&1
cat gpurun_out/g11_leaf.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g11_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g11_tests.log
tail -12 gpurun_out/g11_tests.log
