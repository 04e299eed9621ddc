timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err; tail -3 gpurun_out/bench_r01k.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r01k.json')); x=d['extra']
print('value', d['value'], 'e2e', d['e2e']['value'], 'states', d['e2e']['states_value'], 'roof', d['roofline']['frac'], x['roofline_other_form']['frac'], 'cpu', d['cpu_baseline']['value'])
print(['%.1f'%(1e3*v) for v in x['fused_ms_per_submesh']], ['%.1f'%(1e3*v) for v in x['skin_ms_per_submesh']])
P
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01k.json; cut -c1-200 gpurun_out/bench_ref_r01k.json
