"""The CPU oracle against the committed known answers (tests/golden/known_answers.json, produced by make_golden.py in the
authoring container where the oracle is checked against oracle/_ref) and the reference's README SAH costs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh, random_tris

KA = json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def h32(orc, a):
    return orc.fnv1a(np.ascontiguousarray(a).view(np.uint32).reshape(-1))


def cases():
    out = []
    for name in ("cornellbox", "bunny", "sponza"):
        out.append((name, lambda name=name: load_mesh(name)))
    for key in KA:
        if key.startswith("synth_") and isinstance(KA[key], dict):
            _, kind, n, seed = key.split("_")
            out.append((key, lambda kind=kind, n=int(n), seed=int(seed): random_tris(n, seed, kind)))
    return out


@pytest.mark.parametrize("name,make", cases(), ids=[c[0] for c in cases()])
def test_oracle_reproduces_known_answers(oracle, name, make):
    tris = make()
    if tris is None:
        pytest.skip(f"mesh {name} not staged (oracle/stage_meshes.py)")
    ka = KA[name]
    n = tris.size
    assert n == ka["n"]
    lb = oracle.build_lbvh(tris)
    assert oracle.fnv1a(lb["skeys"], lb["svals"]) == ka["sorted_kv_fnv"]
    assert h32(oracle, lb["nodes"]) == ka["karras_nodes_fnv"]
    assert h32(oracle, lb["wide"]) == ka["lbvh_wide_fnv"] and lb["wide_count"] == ka["lbvh_wide_count"]
    assert np.float32(lb["cost"]) == np.float32(ka["lbvh_cost"])
    sp = oracle.build_lbvh(tris, single_pass=True)
    assert sp["root"] == ka["apetrei_root"] and h32(oracle, sp["nodes"]) == ka["apetrei_nodes_fnv"]
    assert h32(oracle, sp["wide"]) == ka["lbvh_wide_fnv"]
    pl = oracle.build_ploc(tris)
    assert h32(oracle, pl["nodes"]) == ka["ploc_nodes_fnv"] and pl["wide_count"] == ka["ploc_wide_count"]
    assert np.float32(pl["cost"]) == np.float32(ka["ploc_cost"]) and pl["stats"]["iterations"] == ka["ploc_iterations"]
    hp = oracle.build_ploc(tris, hierarchical=True)
    assert h32(oracle, hp["nodes"]) == ka["hploc_nodes_fnv"] and hp["wide_count"] == ka["hploc_wide_count"]
    assert np.float32(hp["cost"]) == np.float32(ka["hploc_cost"]) and hp["stats"]["merge_calls"] == ka["hploc_merge_calls"]
    for t in (lb, sp):
        assert oracle.check_bvh4(t["wide"], t["wide_leaves"], 0, n)
    for t in (pl, hp):
        assert oracle.check_bvh2(t["nodes"], t["leaves"], 0, n) and oracle.check_bvh4(t["wide"], t["wide_leaves"], 0, n)
    if "readme" in ka:  # the reference's only published known answers (README.md:61..207), 6 significant digits
        assert lb["cost"] == pytest.approx(ka["readme"]["lbvh"], rel=1e-5)
        assert pl["cost"] == pytest.approx(ka["readme"]["ploc"], rel=1e-4)
        assert hp["cost"] == pytest.approx(ka["readme"]["hploc"], rel=1e-5)


def test_synth_uniform_v1_is_frozen(oracle):
    t = oracle.synth_uniform(10_000_000, 0x00B20010, first=0, count=4096)
    assert h32(oracle, t["v"].reshape(-1, 9)) == KA["synth_uniform_v1_first4096_fnv"]
    # sharded generation == slices of the whole
    a = oracle.synth_uniform(1000, 7)
    b = oracle.synth_uniform(1000, 7, first=300, count=200)
    assert a[300:500].tobytes() == b.tobytes()


def test_binned_sah_cpu_baseline(oracle):
    tris = load_mesh("cornellbox")
    nodes, cnt = oracle.binned_sah(tris)
    assert cnt == 63 and oracle.check_sah(nodes, tris.size)
    quirk, proper = oracle.cost_binned_sah(nodes)
    assert quirk == pytest.approx(40.4712, rel=1e-5)
    assert proper > 1.0
