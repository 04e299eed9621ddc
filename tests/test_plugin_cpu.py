"""CPU: the C++ plugin is a loadable OGRE plugin with the right entry points and it fails LOUDLY
(an OGRE exception naming the missing device) instead of falling back when there is no GPU."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PLUGIN = os.path.join(ROOT, "ogre_b200", "plugin", "RenderSystem_CUDA.so")
HARNESS = os.path.join(HERE, "harness", "libplugin_harness.so")

need_plugin = pytest.mark.skipif(not (os.path.exists(PLUGIN) and os.path.exists(HARNESS)),
                                 reason="plugin not built (needs the reference tree: __graft_entry__.build())")


@need_plugin
def test_plugin_exports_ogre_entry_points():
    out = subprocess.check_output(["nm", "-D", "--defined-only", PLUGIN], text=True)
    for sym in ("dllStartPlugin", "dllStopPlugin"):
        assert (" T %s" % sym) in out, sym
    # links the C ABI, not a private copy of the kernels
    und = subprocess.check_output(["nm", "-D", "--undefined-only", PLUGIN], text=True)
    for sym in ("ogrecu_init", "ogrecu_draw", "ogrecu_flush", "ogrecu_software_vertex_skinning_host", "ogrecu_buf_map"):
        assert sym in und, sym


@need_plugin
def test_plugin_fails_loudly_without_a_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, os.path.join(HERE, "harness", "run_scene.py"), "--plugin", "cuda", "--out",
                        str(tmp_path / "x.npz")], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)


@need_plugin
def test_harness_with_the_real_tiny_plugin_reproduces_the_golden_image(tmp_path):
    # validates the harness itself: same scene, real Tiny plugin loaded by path -> the committed fixture
    import numpy as np
    out = str(tmp_path / "tiny.npz")
    # The REFERENCE plugin is retried: TinyRenderSystem never initialises DefaultShader::image / mDepthTest / mBlendAdd
    # (ctor, OgreTinyRenderSystem.cpp:27-46) and was seen to die in sample2D on some runs — the oracle's own undefined
    # behaviour, not the harness under validation.  glibc's MALLOC_PERTURB_=255 zero-fills fresh allocations, so those members
    # read as NULL / false instead of heap garbage.
    for attempt in range(4):
        r = subprocess.run([sys.executable, os.path.join(HERE, "harness", "run_scene.py"), "--plugin", "tiny", "--size", "320x180",
                            "--out", out], capture_output=True, text=True, timeout=300, env=dict(os.environ, MALLOC_PERTURB_="255"))
        if r.returncode == 0:
            break
    assert r.returncode == 0, r.stdout + r.stderr
    a = np.load(out)
    g = np.load(os.path.join(HERE, "golden", "ogrehead_320x180.npz"))
    assert np.array_equal(a["colour"], g["colour"])
    assert list(a["stats"]) == [4336, 4]
