for c in 1 2 3; do python tools/bench_configs.py --config c4 --views 256 --contexts $c 2>gpurun_out/c4ctx_$c.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('contexts $c', round(d['views_per_s_device']), round(d['ms_per_view_device'],4))"; tail -2 gpurun_out/c4ctx_$c.err; done
