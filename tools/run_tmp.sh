timeout 600 python -m pytest tests/test_skeleton_gpu.py tests/test_optutil_gpu.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f e2e %.0f states %.0f'%(d['value'], d['e2e']['value'], d['e2e']['states_value']), d['e2e']['states_note'][:140])"
python tools/bench_configs.py --config c3 | cut -c1-200
