#!/bin/sh
TAG=r01q
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_skeleton_gpu.py tests/test_golden_gpu.py -m gpu -q > gpurun_out/pytest_skel_$TAG.log 2>&1; echo "pytest skel rc=$?"; tail -5 gpurun_out/pytest_skel_$TAG.log | cut -c1-300
python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick_r01q.json'))
e=d['e2e']
print('value %.0f e2e %.0f states %.0f ms %.4f states_ms %.4f' % (d['value'], e['value'], e['states_value'], d['ms_per_step'], e['states_ms_per_step']))
print(e['states_note'][:160])
PY
