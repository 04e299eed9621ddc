mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r01e.json
python tools/bench_configs.py --config c1,c3,c4,c5 --views 256 --cpu --out gpurun_out/configs_r01e.jsonl 2>&1 | cut -c1-600
