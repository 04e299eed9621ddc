timeout 900 python -m pytest tests/test_raster_gpu.py tests/test_raster_paths_gpu.py tests/test_golden_gpu.py tests/test_optutil_gpu.py -x -q 2>&1 | tail -3
python tools/bench_configs.py --config c3,c4,c5 --views 256 | cut -c1-1000
