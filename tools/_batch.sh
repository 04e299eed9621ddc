mkdir -p gpurun_out
timeout 120 python tools/lane_bench.py --steps 50 --per-sub --tag parity 2>&1 | tail -1 | tee gpurun_out/batch3.jsonl
timeout 300 python -m pytest tests/test_skin_lane_gpu.py -x -q -m gpu 2>&1 | tail -15
