timeout 900 python -m pytest tests/test_raster_gpu.py tests/test_raster_paths_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -3
python tools/bench_configs.py --config c1,c4,c5 --views 256 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:30], {k:round(v,4) for k,v in d.items() if isinstance(v,float) and ('ms_per' in k or 'views_per_s_device' in k)})"
