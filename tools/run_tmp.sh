#!/bin/sh
# Scratch script for one-off gpurun sessions (edit freely; tools/gpu_session.sh is the full validation session).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_tmp.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_tmp.log | cut -c1-300
