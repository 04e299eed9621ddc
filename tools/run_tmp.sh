timeout 900 python -m pytest tests/test_skin_tile_gpu.py tests/test_full_size_gpu.py -x -q -k "multi or full or fused" 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 200 2>gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms %.4f launches %d e2e %.0f states %.0f roof %.3f'%(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['e2e']['states_value'], d['roofline']['frac']))"; tail -3 gpurun_out/b.err
python bench.py --no-cpu-baseline --steps 100 --form separate 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('separate value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))"
