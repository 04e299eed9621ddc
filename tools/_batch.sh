mkdir -p gpurun_out
for v in default; do
  if [ $v = default ]; then unset OGRECU_LIB; else export OGRECU_LIB=$PWD/build/variants/$v/libogrecu.so; fi
  timeout 120 python tools/lane_bench.py --steps 50 --per-sub --tag $v 2>&1 | tail -1
done | tee gpurun_out/batch2.jsonl
unset OGRECU_LIB
timeout 300 python -m pytest tests/test_skin_lane_gpu.py -x -q -m gpu 2>&1 | tail -3
