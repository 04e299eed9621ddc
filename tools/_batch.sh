mkdir -p gpurun_out
for v in default head lean_ne lean_w13 lean_w11; do
  if [ $v = default ]; then unset OGRECU_LIB; else export OGRECU_LIB=$PWD/build/variants/$v/libogrecu.so; fi
  timeout 120 python tools/lane_bench.py --steps 50 --tag $v 2>&1 | tail -1
done | tee gpurun_out/batch1.jsonl
