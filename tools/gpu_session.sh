#!/bin/sh
# One GPU-box session: smoke, tests, bench (both arms), ncu launch list, ncu full captures of the dominant kernels.
# Usage (under gpurun):  sh tools/gpu_session.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --form separate --no-cpu-baseline > gpurun_out/bench_separate_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "bench separate rc=$?"; cut -c1-400 gpurun_out/bench_separate_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
# launch list (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 42 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu list rc=$?"
# full capture of the dominant kernel: second step, sub-mesh 1 (1888 verts x 10000, 4 weights)
ncu --set full --clock-control none --import-source on -k regex:skin_pair -s 7 -c 1 -o gpurun_out/prof_fused_$TAG -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu full fused rc=$?"
# full capture of the single-call stream kernel (SURVEY 8d roofline case, 4 weights)
ncu --set full --clock-control none --import-source on -k regex:skin_stream -s 3 -c 1 -o gpurun_out/prof_stream_$TAG -f \
    python tools/microbench.py > /dev/null 2>&1; echo "ncu full stream rc=$?"
ls -la gpurun_out | tail -12
