#!/bin/sh
TAG=r01q
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_skin_tile_gpu.py tests/test_full_size_gpu.py tests/test_skeleton_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_pair_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_pair_$TAG.log | cut -c1-400
python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_pair_$TAG.json 2> gpurun_out/bench_pair_$TAG.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_pair_r01q.json'))
e=d['e2e']; x=d['extra']
print('value %.0f e2e %.0f states %.0f ms %.4f roof %.3f other %.3f' % (d['value'], e['value'], e['states_value'], d['ms_per_step'], d['roofline']['frac'], x['roofline_other_form']['frac']))
print('fused per submesh', [round(v,4) for v in x['fused_ms_per_submesh']])
print('skin per submesh', [round(v,4) for v in x['skin_ms_per_submesh']])
PY
