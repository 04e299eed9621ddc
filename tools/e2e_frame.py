#!/usr/bin/env python3
"""One e2e frame of bench.py (skin + render of all instances) for profiling: python tools/e2e_frame.py [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0], "--steps", "3", "--warmup", "3", "--e2e-steps", sys.argv[1] if len(sys.argv) > 1 else "3", "--no-render", "--no-cpu-baseline", "--no-plugin"]
import bench
bench.main()
