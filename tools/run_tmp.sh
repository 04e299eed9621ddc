timeout 600 python -m pytest tests/test_skeleton_gpu.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 200 2>gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f states %.0f %s'%(d['value'], d['e2e']['value'], d['e2e']['states_value'], d['e2e']['states_note'][60:135]))"; tail -3 gpurun_out/b.err
