mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "skin or golden or optutil" 2>&1 | tail -3
for lib in "" build/variants/libogrecu_m4.so build/variants/libogrecu_m6.so; do
  for form in fused separate; do
  OGRECU_LIB=${lib:+$PWD/$lib} python bench.py --no-cpu-baseline --form $form 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['extra']
print('lib=$lib form=$form value=%.0f e2e=%.0f fused=%.0f separate=%.0f skin_only=%.0f morph_only=%.0f roof=%.3f other=%.3f' % (d['value'], d['e2e']['value'], x['fused_Mvert_per_s'], x['separate_Mvert_per_s'], x['skin_only_Mvert_per_s'], x['morph_only_Mvert_per_s'], d['roofline']['frac'], x['roofline_other_form']['frac']))"
  done
done | tee gpurun_out/sweep_r01c.txt
ncu --set full --clock-control none --import-source on -k regex:skin_tile -s 8 -c 1 -o gpurun_out/prof_fused_r01c -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"
