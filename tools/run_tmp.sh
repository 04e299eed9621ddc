#!/bin/sh
TAG=r01q
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest all rc=$?"; tail -25 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
for sync in 0 1; do
  echo "cuda Sinbad optutil_sync=$sync"; OGRECU_OPTUTIL_SYNC=$sync python tests/harness/run_scene.py --plugin cuda --mesh Sinbad.mesh --size 256x256 --cam-z 30 --anim-time 300 --out /tmp/c.npz | grep anim_frame
done
