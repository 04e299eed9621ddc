timeout 900 python -m pytest tests/test_skin_tile_gpu.py -x -q 2>&1 | tail -3
OGRECU_LIB=$PWD/ogre_b200/libogrecu_t128_c6.so timeout 900 python -m pytest tests/test_skin_tile_gpu.py -x -q 2>&1 | tail -3
for lib in libogrecu.so libogrecu_t128_c6.so; do
OGRECU_LIB=$PWD/ogre_b200/$lib python bench.py --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['extra']
print('$lib value=%.0f fused=%.0f separate=%.0f skin_only=%.0f roof=%.3f (%.3f ms) other=%.3f (%.3f ms)' % (d['value'], x['fused_Mvert_per_s'], x['separate_Mvert_per_s'], x['skin_only_Mvert_per_s'], d['roofline']['frac'], d['roofline']['launch_ms'], x['roofline_other_form']['frac'], x['roofline_other_form']['launch_ms']))"
done
