#!/usr/bin/env python3
"""Is the REFERENCE's own frame a function of the host?  Runs the unmodified Tiny plugin (oracle/_ref) on this machine for the
scenes whose frames were dumped on the build host (tests/golden/skel_frames.npz) and reports every pixel that differs, for
several fill patterns of fresh heap memory (glibc MALLOC_PERTURB_: an uninitialised read shows up as a dependence on it).
usage: python tools/tiny_host_check.py"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "tests", "harness", "run_scene.py")
G = np.load(os.path.join(ROOT, "tests", "golden", "skel_frames.npz"))
SCENES = [("Sinbad.mesh", ["--mesh", "Sinbad.mesh", "--size", "256x256", "--cam-z", "30", "--skel", "0.3,1.7"], "Sinbad.mesh"),
          ("jaiqua.mesh", ["--mesh", "jaiqua.mesh", "--size", "256x256", "--cam-z", "30", "--skel", "0.3,1.7"], "jaiqua.mesh"),
          ("ogrehead", ["--mesh", "ogrehead.mesh", "--size", "400x300", "--cam-z", "90"], "ogrehead_400x300_z90")]
cpu = ""
try:
    cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    pass
out = {"cpu": cpu, "scenes": {}}
for name, args, key in SCENES:
    res = {}
    for perturb in ("0", "85", "255", None, "mxcsr=0x1f80", "mxcsr=0x9fc0"):
        env = dict(os.environ)
        if perturb and perturb.startswith("mxcsr="):
            env["PH_MXCSR"] = perturb.split("=")[1]
            env["MALLOC_PERTURB_"] = "255"
        elif perturb is None:
            env.pop("MALLOC_PERTURB_", None)
        else:
            env["MALLOC_PERTURB_"] = perturb
        if False:
            pass
        with tempfile.TemporaryDirectory() as d:
            o = os.path.join(d, "t.npz")
            r = subprocess.run([sys.executable, RUN, "--plugin", "tiny", "--out", o] + args, capture_output=True, text=True, env=env, timeout=600)
            if r.returncode != 0:
                res[str(perturb)] = {"crashed": r.returncode, "stderr": r.stderr[-300:]}
                continue
            z = np.load(o)
            c = z["colour"]
            mx = [hex(int(z["mxcsr_at_start"])), hex(int(z["mxcsr_after_open"]))]
        bad = np.argwhere((c != G[key]).any(axis=2))
        res[str(perturb)] = {"mxcsr": mx, "differing_pixels": int(len(bad)), "where": bad[:8].tolist(),
                             "values": [[c[y, x].tolist(), G[key][y, x].tolist()] for y, x in bad[:4]]}
    out["scenes"][name] = res
print(json.dumps(out))
