"""The C++ host API (reference class names) end to end on the GPU: b2bvh_demo builds with each builder object, traces,
prints the reference's perf table, and its m_cost equals the oracle's."""
import os
import subprocess

import numpy as np
import pytest

from conftest import MESH_DIR, GOLDEN, ROOT, load_mesh
from b2bvh import capi

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "hip-bvh-construction_b200", "b2bvh_demo")


@pytest.mark.parametrize("which,hier", [("twopass", None), ("singlepass", None), ("ploc", False), ("hploc", True)])
def test_demo_matches_oracle_cost(oracle, which, hier):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hip-bvh-construction_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    mesh = "bunny" if load_mesh("bunny") is not None else "cornellbox"
    path = os.path.join(MESH_DIR if mesh == "bunny" else GOLDEN, mesh + ".tri")
    tris = load_mesh(mesh)
    o = oracle.build_lbvh(tris) if hier is None else oracle.build_ploc(tris, hierarchical=hier)
    env = dict(os.environ, B2BVH_LIB=capi.LIB_PATH)
    r = subprocess.run([EXE, which, path, repr(float(np.float32(o["cost"])))], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for token in ("Executing on", "CalculateCentroidExtentsTime", "SortingTime", "BvhBuildTime", "CollapseTime", "Bvh Cost", "Total Time"):
        assert token in r.stdout
    assert f"wide nodes : {o['wide_count']}" in r.stdout


def test_demo_batched_builder():
    """The USE_BATCHED_BUILDER branch of main.cpp:38-52: 4096 copies of the cornell box through BatchedBvhBuilder."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hip-bvh-construction_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    env = dict(os.environ, B2BVH_LIB=capi.LIB_PATH)
    r = subprocess.run([EXE, "batched", os.path.join(GOLDEN, "cornellbox.tri")], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BatchSize : 4096" in r.stdout and "BvhBuildTime" in r.stdout and f"nodes : {4096 * 31}" in r.stdout


def test_demo_twopass_with_prim_splitting(oracle):
    """TwoPassLbvh::m_saMax = the reference's USE_PRIM_SPLITTING build (saMax = 10 there): reference count and cost equal the oracle's."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hip-bvh-construction_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    tris = load_mesh("cornellbox")
    o = oracle.build_lbvh(tris, split_sa_max=10.0)
    env = dict(os.environ, B2BVH_LIB=capi.LIB_PATH)
    r = subprocess.run([EXE, "twopass-split:10", os.path.join(GOLDEN, "cornellbox.tri"), repr(float(np.float32(o["cost"])))], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"references : {o['refs'].size} of {tris.size} triangles" in r.stdout and f"wide nodes : {o['wide_count']}" in r.stdout

