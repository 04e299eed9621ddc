#!/bin/sh
TAG=r01q
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hwbuffer_gpu.py tests/test_plugin_gpu.py tests/test_optutil_gpu.py tests/test_pose_blend_gpu.py tests/test_shadow_gpu.py -m gpu -q > gpurun_out/pytest_new2_$TAG.log 2>&1; echo "pytest new rc=$?"; tail -30 gpurun_out/pytest_new2_$TAG.log | cut -c1-300
for mesh in Sinbad.mesh jaiqua.mesh; do
for sync in 0 1 0 1; do
  echo "cuda $mesh optutil_sync=$sync"; OGRECU_OPTUTIL_SYNC=$sync python tests/harness/run_scene.py --plugin cuda --mesh $mesh --size 256x256 --cam-z 30 --anim-time 300 --out /tmp/c.npz | grep anim_frame
done
echo "tiny $mesh"; MALLOC_PERTURB_=255 python tests/harness/run_scene.py --plugin tiny --mesh $mesh --size 256x256 --cam-z 30 --anim-time 100 --out /tmp/t.npz | grep anim_frame
done 2>&1 | tee gpurun_out/plugin_anim_$TAG.txt
