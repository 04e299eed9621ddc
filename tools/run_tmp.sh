for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/bench_configs.py --config c5 > gpurun_out/c5_sl_n$n.out 2> gpurun_out/c5_sl_n$n.err; echo rc=$?
grep '^{' gpurun_out/c5_sl_n$n.out > gpurun_out/configs_c5_n${n}_r01n.jsonl
python - <<P
import json
d=json.loads(open('gpurun_out/configs_c5_n${n}_r01n.jsonl').read())
print($n,'strips', round(d['ms_per_frame_render_only'],3), round(d['ms_per_frame_with_composite'],3), d['identical_to_single_gpu_frame'])
print($n,'bands ', round(d['interleaved']['ms_per_frame_render_only'],3), round(d['interleaved']['ms_per_frame_with_composite'],3), d['interleaved']['identical_to_single_gpu_frame'])
print($n,"sortlast", round(d["sort_last"]["ms_per_frame_render_only"],3), round(d["sort_last"]["ms_per_frame_with_composite"],3), round(d["sort_last"]["ms_composite_only"],3), d['sort_last']['identical_to_single_gpu_frame'])
P
done
