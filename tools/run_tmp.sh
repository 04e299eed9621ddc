python tools/bench_configs.py --config c3 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['ms_per_frame'], json.dumps(d['shadow_volume_indices_8']))"
