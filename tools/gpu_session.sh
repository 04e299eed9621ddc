#!/bin/sh
# One GPU-box session: smoke, tests, bench (both arms), ncu launch list of an e2e frame, ncu full captures of the dominant kernels.
# Usage (under gpurun):  [SESSION_SETUP_COUNT=1] sh tools/gpu_session.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
# launch list of e2e frames: palettes, repack + lane kernel, the render of all instances (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_e2e_$TAG.csv \
    python tools/e2e_frame.py 3 > /dev/null 2>&1; echo "ncu list rc=$?"
# full captures: the lane kernel on the bench step, the raster set-up count pass on the crowd frame
ncu --set full --clock-control none --import-source on -k regex:skin_lane -s 8 -c 1 -o gpurun_out/prof_lane_$TAG -f \
    python tools/e2e_frame.py 2 > /dev/null 2>&1; echo "ncu full lane rc=$?"
if [ -n "$SESSION_SETUP_COUNT" ]; then
ncu --set full --clock-control none --import-source on -k regex:setup_count -s 6 -c 1 -o gpurun_out/prof_setup_count_$TAG -f \
    python tools/e2e_frame.py 2 > /dev/null 2>&1; echo "ncu full setup_count rc=$?"
fi
ls -la gpurun_out | tail -12
