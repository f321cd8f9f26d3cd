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


def mesh_file(mesh, tmp_path):
    """Path of the raw triangle file of a mesh: the staged copy, or the committed fixture unpacked into tmp_path."""
    for d in (GOLDEN, MESH_DIR):
        p = os.path.join(d, mesh + ".tri")
        if os.path.exists(p):
            return p
    p = str(tmp_path / (mesh + ".tri"))
    load_mesh(mesh)["v"].reshape(-1, 9).astype(np.float32).tofile(p)
    return p


@pytest.mark.parametrize("which,hier", [("twopass", None), ("singlepass", None), ("ploc", False), ("hploc", True)])
def test_demo_matches_oracle_cost(oracle, which, hier, tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hip-bvh-construction_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    mesh = "bunny" if load_mesh("bunny") is not None else "cornellbox"
    path = mesh_file(mesh, tmp_path)
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



def test_demo_sharded_build(oracle, tmp_path):
    """ShardedLbvh over b2bvh_build_sharded (one host thread, G contexts): the per-shard wide-node counts and the scene box are the
    oracle's sharded build's (build_sharded)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hip-bvh-construction_b200", "host"), "all"], stdout=subprocess.DEVNULL)
    tris = load_mesh("bunny") if load_mesh("bunny") is not None else load_mesh("cornellbox")
    path = mesh_file("bunny" if tris.size > 1000 else "cornellbox", tmp_path)
    env = dict(os.environ, B2BVH_LIB=capi.LIB_PATH)
    r = subprocess.run([EXE, "sharded:3", path], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    scene, shards, top = oracle.build_sharded(tris, 3, single_pass=True)
    for g, s in enumerate(shards):
        assert f"Shard {g} : {s['skeys'].size} primitives" in r.stdout and f"wide nodes {s['wide_count']}" in r.stdout
    assert "shards : 3  top-level nodes : 5" in r.stdout
