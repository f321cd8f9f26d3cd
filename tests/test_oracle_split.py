"""Early split clipping: the oracle's restatement of Utility::doEarlySplitClipping (Utility.cpp:456-538) against
the reference's own function (oracle/_ref/libref_utility.so, when /root/reference was present at build time) and against the
committed known answers (tests/golden/split_known_answers.json, tests/golden/make_golden_split.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_mesh, random_tris

KA = json.load(open(os.path.join(GOLDEN, "split_known_answers.json")))
CASES = [k for k in KA if not k.startswith("_")]


def make_case(key):
    kind, n, seed, sa = key.rsplit("_", 3)
    tris = load_mesh(kind) if n == "None" else random_tris(int(n), int(seed), kind)
    return tris, float(sa)


def h32(orc, a):
    return orc.fnv1a(np.ascontiguousarray(a).view(np.uint32).reshape(-1))


@pytest.mark.parametrize("key", CASES)
def test_split_known_answers(oracle, key):
    tris, sa = make_case(key)
    ka = KA[key]
    refs = oracle.early_split(tris, sa)
    assert refs.size == ka["n_refs"] and h32(oracle, refs) == ka["refs_fnv"]
    b = oracle.build_lbvh(tris, split_sa_max=sa)
    assert h32(oracle, b["nodes"]) == ka["nodes_fnv"] and h32(oracle, b["wide"]) == ka["wide_fnv"]
    assert h32(oracle, b["wide_leaves"]) == ka["wide_leaves_fnv"] and b["wide_count"] == ka["wide_count"]
    assert np.float32(b["cost"]) == np.float32(ka["cost"])


def test_split_properties(oracle):
    """Size-independent properties: every reference fits saMax, the fragments of a triangle tile its box exactly (volumes add up,
    union = the triangle's box), references of one generation keep the queue order, leaves of the tree name triangles."""
    tris = random_tris(2000, 31, "uniform")
    sa = 1.0
    refs = oracle.early_split(tris, sa)
    _, boxes, _ = oracle.primrefs(tris)
    e = (refs["mx"] - refs["mn"]).astype(np.float32)
    area = np.float32(2) * ((e[:, 0] * e[:, 1] + e[:, 0] * e[:, 2]) + e[:, 1] * e[:, 2])
    assert (area <= np.float32(sa)).all()
    prim = refs["primIdx"]
    assert set(prim.tolist()) == set(range(tris.size))
    for t in (0, 7, 1999):
        m = prim == t
        assert np.array_equal(refs["mn"][m].min(axis=0), boxes["mn"][t]) and np.array_equal(refs["mx"][m].max(axis=0), boxes["mx"][t])
        vol = np.prod((refs["mx"][m] - refs["mn"][m]).astype(np.float64), axis=1).sum()
        assert abs(vol - np.prod((boxes["mx"][t] - boxes["mn"][t]).astype(np.float64))) <= 1e-5 * max(vol, 1e-30)
    b = oracle.build_lbvh(tris, split_sa_max=sa)
    n = refs.size
    leaves = b["nodes"][n - 1:]
    assert np.array_equal(np.sort(leaves["left"]), np.sort(prim)) and (leaves["left"] < tris.size).all()
    assert np.array_equal(leaves["left"], prim[b["svals"]])
    assert oracle.check_root_aabb(b["nodes"], 0, n)


def test_split_off_is_identity(oracle):
    tris = random_tris(500, 32, "uniform")
    refs = oracle.early_split(tris, 3.0e38)
    base, _, _ = oracle.primrefs(tris)
    assert refs.tobytes() == base.tobytes()


def test_split_nonterminating_is_reported(oracle):
    tris = random_tris(10, 33, "uniform")
    tris["v"][3, 0, 0] = np.inf
    with pytest.raises(ValueError):
        oracle.early_split(tris, 1.0)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_utility.so")), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind,n,seed,sa", [("uniform", 4000, 41, 6.0), ("uniform", 1500, 42, 0.7), ("clustered", 3000, 43, 1e-4), ("anisotropic", 2500, 44, 0.003),
                                            ("flat", 2000, 45, 0.02), ("duplicate", 100, 46, 0.5)])
def test_split_matches_reference_function(oracle, kind, n, seed, sa):
    import ref
    tris = random_tris(n, seed, kind)
    a, b = oracle.early_split(tris, sa), ref.early_split_sa(tris, sa)
    assert a.size == b.size and a.tobytes() == b.tobytes()
