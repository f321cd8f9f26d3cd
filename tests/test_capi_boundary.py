"""The drop-in boundary without a GPU: libb2bvh.so loads, exports every symbol include/b2bvh.h declares, the struct
layouts the Python mirror uses agree with the header, and the product refuses to run without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from b2bvh import capi, types as T


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2bvh.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2bvh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = header_symbols()
    assert declared, "no declarations parsed"
    assert sorted(capi.SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.b2bvh_abi_version() == capi.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "b2bvh.h")).read()
    assert f"#define B2BVH_ABI_VERSION {capi.ABI_VERSION}u" in hdr


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b2bvh.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b2bvh_tree), sizeof(b2bvh_build_opts),'
                   ' offsetof(b2bvh_tree, d_bvhNodes), offsetof(b2bvh_tree, stage_ms), offsetof(b2bvh_tree, n_launches), offsetof(b2bvh_tree, d_primRefIdx),'
                   ' offsetof(b2bvh_build_opts, split_sa_max));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])  # the header is plain C
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(capi.Tree), C.sizeof(capi.BuildOpts), capi.Tree.d_bvhNodes.offset, capi.Tree.stage_ms.offset, capi.Tree.n_launches.offset,
                   capi.Tree.d_primRefIdx.offset, capi.BuildOpts.split_sa_max.offset]
    assert T.TRIANGLE.itemsize == 64 and T.BVH2_NODE.itemsize == 32 and T.BVH4_NODE.itemsize == 128 and T.PRIM_REF.itemsize == 28


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.B2bvhError, match="no CUDA device"):
        capi.Context(0)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "hip-bvh-construction_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(d, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "oracle/" not in text.replace("as in oracle", ""), os.path.join(d, f)


def test_cpp_host_layer_builds_and_fails_loudly_without_gpu():
    """The reference-compatible C++ classes (host/BvhConstruction.h) compile, bind libb2bvh.so with dlopen, and refuse to run
    without a CUDA device (no CPU fallback)."""
    import torch
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hip-bvh-construction_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "hip-bvh-construction_b200", "b2bvh_demo")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    env = dict(os.environ, B2BVH_LIB=capi.LIB_PATH)
    r = subprocess.run([exe, "twopass", "synth:100"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
