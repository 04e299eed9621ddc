#!/bin/sh
# One GPU-box session: tests, bench, ncu launch list, ncu full capture of the dominant kernel.
# Usage (under gpurun):  sh tools/gpu_session.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --form separate --no-cpu-baseline > gpurun_out/bench_separate_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "bench separate rc=$?"; cat gpurun_out/bench_separate_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
# A/B sweep of the kernel-selection knobs (short runs)
for cfg in "0 0" "3 0" "4 0"; do
  set -- $cfg
  python bench.py --no-cpu-baseline --steps 50 --skin-mode $1 --skin-ipc $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['extra']
print('mode=$1 ipc=$2 value=%.0f fused=%.0f separate=%.0f skin_only=%.0f roof=%.3f other=%.3f' % (d['value'], x['fused_Mvert_per_s'], x['separate_Mvert_per_s'], x['skin_only_Mvert_per_s'], d['roofline']['frac'], x['roofline_other_form']['frac']))"
done | tee gpurun_out/sweep_$TAG.txt
# launch list (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 42 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu list rc=$?"
# full capture of the dominant kernel: 9th skin_tile launch = second step, sub-mesh 1 (1888 verts x 10000, 4 weights)
ncu --set full --clock-control none --import-source on -k regex:skin_ -s 7 -c 1 -o gpurun_out/prof_fused_$TAG -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu full fused rc=$?"
ncu --set full --clock-control none --import-source on -k regex:skin_ -s 7 -c 1 -o gpurun_out/prof_skin_$TAG -f \
    python bench.py --form separate --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu full separate rc=$?"
ls -la gpurun_out | tail -14
