mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_configs.py --config c1,c4,c5 --views 256 --out gpurun_out/configs_vis.jsonl 2>&1 | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c5.csv python tools/bench_configs.py --config c5 > /dev/null 2>&1
