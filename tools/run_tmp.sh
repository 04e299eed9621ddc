mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for p in tiny cuda; do python tests/harness/run_scene.py --plugin $p --mesh ogrehead.mesh --size 1280x720 --time 20 --out gpurun_out/c1_$p.npz; python -c "
import numpy as np; a=np.load('gpurun_out/c1_$p.npz'); print('$p frame_seconds', float(a['frame_seconds']))"; done
