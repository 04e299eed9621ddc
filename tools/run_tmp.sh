timeout 900 python -m pytest tests/test_skin_tile_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['extra']
print('value %.0f e2e %.0f states %.0f roof %.3f other %.3f'%(d['value'], d['e2e']['value'], d['e2e']['states_value'], d['roofline']['frac'], x['roofline_other_form']['frac']), ['%.1f'%(1e3*v) for v in x['fused_ms_per_submesh']], ['%.1f'%(1e3*v) for v in x['skin_ms_per_submesh']])"
