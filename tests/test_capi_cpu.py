"""CPU: the C-ABI library loads, exports every symbol include/ogrecu.h declares, and fails
loudly (no fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ogrecu.h")).read()
    return sorted(set(re.findall(r"OGRECU_API\s+[\w\s\*]+?\b(ogrecu_\w+)\s*\(", text)))


def test_header_symbols_are_all_exported_and_bound():
    from ogre_b200 import capi
    names = declared_symbols()
    assert len(names) >= 50
    lib = capi.lib()
    for n in names:
        assert hasattr(lib, n), "libogrecu.so does not export %s" % n
        assert n in capi.SIGNATURES, "ogre_b200/capi.py does not bind %s" % n
    assert sorted(capi.SIGNATURES) == names
    assert lib.ogrecu_abi_version() == 6


def test_struct_layouts_match_the_header(tmp_path):
    # ask the C compiler for the layouts of include/ogrecu.h and compare with the ctypes mirrors
    import subprocess
    from ogre_b200 import capi
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ogrecu.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ogrecu_skin_job), sizeof(ogrecu_draw_desc),'
        ' sizeof(ogrecu_raster_stats), offsetof(ogrecu_draw_desc, texture), offsetof(ogrecu_draw_desc, mvp),'
        ' offsetof(ogrecu_draw_desc, ambient), offsetof(ogrecu_skin_job, num_vertices), sizeof(ogrecu_anim_job),'
        ' offsetof(ogrecu_anim_job, dst_instance_stride_bytes), sizeof(ogrecu_skeleton_desc), sizeof(ogrecu_animation_desc),'
        ' sizeof(ogrecu_draw_state), offsetof(ogrecu_draw_state, scissor), offsetof(ogrecu_draw_state, colours));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(capi.SkinJob), C.sizeof(capi.DrawDesc), C.sizeof(capi.RasterStats), capi.DrawDesc.texture.offset,
                   capi.DrawDesc.mvp.offset, capi.DrawDesc.ambient.offset, capi.SkinJob.num_vertices.offset, C.sizeof(capi.AnimJob),
                   capi.AnimJob.dst_instance_stride_bytes.offset, C.sizeof(capi.SkeletonDesc), C.sizeof(capi.AnimationDesc),
                   C.sizeof(capi.DrawState), capi.DrawState.scissor.offset, capi.DrawState.colours.offset]


def test_no_device_is_a_loud_failure_not_a_fallback():
    import ogre_b200
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(ogre_b200.OgrecuError) as e:
        ogre_b200.Context(0)
    assert e.value.status == 1      # OGRECU_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under ogre_b200/ may reference it
    pkg = os.path.join(ROOT, "ogre_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "ogre_oracle" not in src and "libref_driver" not in src, f


def test_header_is_plain_c(tmp_path):
    # the drop-in boundary is a C ABI: include/ogrecu.h must compile as C99 (and as C++11) with nothing but itself
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "ogrecu.h"\nint main(void){return (int)sizeof(ogrecu_anim_job) * 0;}\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "h.o")])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-c", str(src), "-o", str(tmp_path / "h2.o")])
