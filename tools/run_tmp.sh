timeout 600 python -m pytest tests/test_raster_gpu.py -x -q -k "packed32" 2>&1 | tail -3
python tools/bench_configs.py --config c1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:round(v,4) for k,v in d.items() if isinstance(v,float)})"
