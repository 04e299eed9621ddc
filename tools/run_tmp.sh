mkdir -p gpurun_out
python tools/bench_configs.py --config c1,c3 --cpu --out gpurun_out/configs_c1c3.jsonl 2>&1 | tail -4
python tools/bench_configs.py --config c4 --views 64 --cpu --out gpurun_out/configs_c4_64.jsonl 2>&1 | tail -3
python tools/bench_configs.py --config c5 --c5-tris 2000000 --out gpurun_out/configs_c5_2M.jsonl 2>&1 | tail -3
