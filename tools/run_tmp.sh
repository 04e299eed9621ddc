mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
