import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
MESH_DIR = os.path.join(ROOT, "oracle", "_ref", "meshes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_mesh(name):
    """Triangles of a reference mesh in the reference loader's order: the committed golden fixture or the staged copy under
    oracle/_ref/meshes (oracle/stage_meshes.py).  Returns None when it is not available."""
    from b2bvh import types as T
    for d in (GOLDEN, MESH_DIR):
        p = os.path.join(d, name + ".tri")
        if os.path.exists(p):
            return T.triangles_from_array(np.fromfile(p, dtype=np.float32).reshape(-1, 9))
    # committed fixture: the same bytes, xz-compressed (tests/golden/<name>.tri.xz, written from oracle/stage_meshes.py's output), so a
    # clone without /root/reference still runs every mesh test instead of skipping it
    p = os.path.join(GOLDEN, name + ".tri.xz")
    if os.path.exists(p):
        import lzma
        return T.triangles_from_array(np.frombuffer(lzma.open(p).read(), dtype=np.float32).reshape(-1, 9).copy())
    return None


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ctx():
    from b2bvh import capi
    c = capi.Context(0)
    yield c
    c.close()


def random_tris(n, seed, kind="uniform"):
    """Seeded synthetic triangle soups for parity tests (float32, reproducible)."""
    from b2bvh import types as T
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        c = rng.uniform(-100, 100, size=(n, 1, 3))
        v = c + rng.uniform(-1, 1, size=(n, 3, 3))
    elif kind == "clustered":  # many duplicate Morton codes
        c = rng.integers(0, 8, size=(n, 1, 3)).astype(np.float64) * 10
        v = c + rng.uniform(-0.01, 0.01, size=(n, 3, 3))
    elif kind == "flat":  # zero extent along z: exercises the 2-D Morton branch
        c = rng.uniform(-5, 5, size=(n, 1, 3))
        v = c + rng.uniform(-0.1, 0.1, size=(n, 3, 3))
        v[:, :, 2] = 1.5
    elif kind == "duplicate":  # every triangle identical: all keys equal, order decided by the index tie-break
        v = np.tile(rng.uniform(-1, 1, size=(1, 3, 3)), (n, 1, 1))
        v[0] += 3.0  # keep the scene non-degenerate
    elif kind == "anisotropic":  # long thin scene: extended Morton pre-bits
        c = rng.uniform(-1, 1, size=(n, 1, 3)) * np.array([1000.0, 10.0, 1.0])
        v = c + rng.uniform(-0.05, 0.05, size=(n, 3, 3))
    else:
        raise ValueError(kind)
    return T.triangles_from_array(v.astype(np.float32))
