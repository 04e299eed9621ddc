timeout 600 python -m pytest tests/test_raster_gpu.py -x -q -k sort_last 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/bench_configs.py --config c5 > gpurun_out/c5_sl_n2.out 2> gpurun_out/c5_sl_n2.err; echo rc=$?
grep '^{' gpurun_out/c5_sl_n2.out > gpurun_out/configs_c5_n2_r01n.jsonl
python - <<P
import json
d=json.loads(open('gpurun_out/configs_c5_n2_r01n.jsonl').read())
print('strips', round(d['ms_per_frame_render_only'],3), round(d['ms_per_frame_with_composite'],3), d['identical_to_single_gpu_frame'])
print('sortlast', round(d['sort_last']['ms_per_frame_render_only'],3), round(d['sort_last']['ms_per_frame_with_composite'],3), round(d['sort_last']['ms_composite_only'],3), d['sort_last']['identical_to_single_gpu_frame'])
P
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/c5_sl_n2.err | tail -8
