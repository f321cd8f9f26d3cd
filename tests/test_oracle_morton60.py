"""60-bit Morton variant, CPU side: known answers of the code definition and its relation to the pinned 10-bit interleave."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh, random_tris

KA = json.load(open(os.path.join(GOLDEN, "morton60_known_answers.json")))


@pytest.mark.parametrize("key", [k for k in KA if not k.startswith("_")])
def test_morton60_definition_is_frozen(oracle, key):
    """The committed known answers of the 60-bit variant (tests/golden/make_golden_morton60.py): the definition has no reference to
    be pinned to, so it is frozen instead."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk60", os.path.join(GOLDEN, "make_golden_morton60.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    kind, n, seed = key.rsplit("_", 2)
    tris = load_mesh(kind) if n == "None" else random_tris(int(n), int(seed), kind)
    got = mk.summarize(tris)
    for f, v in KA[key].items():
        assert got[f] == (pytest.approx(v, rel=0, abs=0) if isinstance(v, float) else v), f


def test_morton60_known_answers(oracle):
    assert oracle.morton60_point([0, 0, 0]) == 0 and oracle.morton60_point([1, 1, 1]) == (1 << 60) - 1
    assert oracle.morton60_point([0.5, 0, 0]) == 1 << 59 and oracle.morton60_point([0, 0.5, 0]) == 1 << 58 and oracle.morton60_point([0, 0, 0.5]) == 1 << 57
    step = 1.0 / (1 << 20)
    assert oracle.morton60_point([step, 0, 0]) == 4 and oracle.morton60_point([0, step, 0]) == 2 and oracle.morton60_point([0, 0, step]) == 1
    assert oracle.morton60_point([float("nan"), -3.0, 7.0]) == oracle.morton60_point([0, 0, 1])


def test_morton60_upper_half_is_the_plain_30_bit_code(oracle):
    """code >> 30 == computeMortonCode of the same point (the 10-bit-per-axis grid the reference's batched builder uses, pinned in
    test_oracle_batched.py) whenever p*2^20 truncated >> 10 == p*2^10 truncated, i.e. for points on the 2^-20 grid."""
    rng = np.random.default_rng(91)
    q = rng.integers(0, 1 << 20, size=(2000, 3))
    for row in q:
        p = (row.astype(np.float64) / (1 << 20)).astype(np.float32)  # exact in float32
        assert oracle.morton60_point(p) >> 30 == oracle.morton_plain(p)


def test_morton60_build_is_a_refinement(oracle):
    """Same triangles, 30-bit plain order vs 60-bit order: the 60-bit sorted sequence is sorted by its upper 30 bits too, and the
    tree over it is a valid LBVH (root box = scene box, both numberings collapse to the same Bvh4)."""
    tris = random_tris(4000, 92, "clustered")
    a = oracle.build_lbvh(tris, morton_bits=60)
    b = oracle.build_lbvh(tris, single_pass=True, morton_bits=60)
    hi = a["skeys"] >> np.uint64(30)
    assert (hi[1:] >= hi[:-1]).all()
    assert oracle.check_root_aabb(a["nodes"], 0, tris.size)
    assert a["wide"].tobytes() == b["wide"].tobytes() and a["cost"] == b["cost"]
