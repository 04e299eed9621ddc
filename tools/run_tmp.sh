timeout 1200 python -m pytest tests/test_full_size_gpu.py tests/test_raster_gpu.py -x -q 2>&1 | tail -12
