#!/bin/sh
TAG=r01q
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_skin_tile_gpu.py tests/test_optutil_gpu.py -m gpu -q -x > gpurun_out/pytest_stream3_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_stream3_$TAG.log | cut -c1-400
for c in 0 1 2 3; do echo "copies shift $c"; OGRECU_STREAM_COPIES=$c timeout 200 python tools/microbench.py 2>&1 | head -3; done | tee gpurun_out/microbench4_$TAG.txt
echo default; timeout 200 python tools/microbench.py 2>&1 | head -3 | tee -a gpurun_out/microbench4_$TAG.txt
