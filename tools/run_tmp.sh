for c in 8 16 32 64; do python tools/bench_configs.py --config c4 --views 256 --chunk $c | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('chunk $c', round(d['views_per_s_device']), round(d['ms_per_view_device'],4))"; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'states', round(d['e2e']['states_value']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | grep '^{' | cut -c1-160
