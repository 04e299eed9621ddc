for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/bench_n${n}_r01o.json; python -c "
import json
d=json.load(open('gpurun_out/bench_n${n}_r01o.json')); print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'states', round(d['e2e']['states_value']), d['ms_per_step'])"
done
