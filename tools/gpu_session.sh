#!/bin/sh
# One GPU-box session: tests, bench, ncu launch list, ncu full capture of the dominant kernel.
# Usage (under gpurun):  sh tools/gpu_session.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json | cut -c1-300
kill $SMI
# launch list (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 42 -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu list rc=$?"
# full capture of the dominant kernel: 9th skin launch = step 2, sub-mesh 1 (1888 verts x 10000, 4 weights)
ncu --set full --clock-control none --import-source on -k regex:skin_instanced -s 8 -c 1 -o gpurun_out/prof_skin_$TAG -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
