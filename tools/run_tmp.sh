#!/bin/sh
mkdir -p gpurun_out/dbg
A="--mesh Sinbad.mesh --size 256x256 --cam-z 30 --skel 0.3,1.7"
python tests/harness/run_scene.py --plugin cuda $A --out gpurun_out/dbg/cuda1.npz
python tests/harness/run_scene.py --plugin cuda $A --out gpurun_out/dbg/cuda2.npz
python tests/harness/run_scene.py --plugin tiny $A --out gpurun_out/dbg/tiny_plain.npz
MALLOC_PERTURB_=255 python tests/harness/run_scene.py --plugin tiny $A --out gpurun_out/dbg/tiny_perturb.npz
python tests/harness/run_scene.py --plugin cuda --mesh Sinbad.mesh --size 256x256 --cam-z 30 --out gpurun_out/dbg/cuda_noskel.npz
python tests/harness/run_scene.py --plugin tiny --mesh Sinbad.mesh --size 256x256 --cam-z 30 --out gpurun_out/dbg/tiny_noskel.npz
