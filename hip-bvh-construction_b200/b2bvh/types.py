"""numpy dtypes that are byte-for-byte the POD layouts of include/b2bvh_types.h
(reference: src/Common.h, sizes/offsets in SURVEY.md Appendix A)."""
import numpy as np

INVALID = 0xFFFFFFFF
FLT_MAX = np.float32(3.402823466e+38)

TRIANGLE = np.dtype({"names": ["v"], "formats": [("<f4", (3, 3))], "offsets": [0], "itemsize": 64})
AABB = np.dtype({"names": ["mn", "mx"], "formats": [("<f4", (3,)), ("<f4", (3,))], "offsets": [0, 12], "itemsize": 24})
BVH2_NODE = np.dtype({"names": ["left", "right", "mn", "mx"], "formats": ["<u4", "<u4", ("<f4", (3,)), ("<f4", (3,))],
                      "offsets": [0, 4, 8, 20], "itemsize": 32})
BVH4_NODE = np.dtype({"names": ["aabb", "child", "parent", "childCount", "pad"],
                      "formats": [("<f4", (4, 6)), ("<u4", (4,)), "<u4", "<u4", ("<u4", (2,))],
                      "offsets": [0, 96, 112, 116, 120], "itemsize": 128})
SAH_NODE = np.dtype({"names": ["mn", "mx", "firstChild", "primCount"], "formats": [("<f4", (3,)), ("<f4", (3,)), "<u4", "<u4"],
                     "offsets": [0, 12, 24, 28], "itemsize": 32})
PRIM_REF = np.dtype({"names": ["primIdx", "mn", "mx"], "formats": ["<u4", ("<f4", (3,)), ("<f4", (3,))], "offsets": [0, 4, 16], "itemsize": 28})
PRIM_NODE = np.dtype({"names": ["primIdx", "parent"], "formats": ["<u4", "<u4"], "offsets": [0, 4], "itemsize": 8})
RAY = np.dtype({"names": ["origin", "direction", "tMin", "tMax"], "formats": [("<f4", (3,)), ("<f4", (3,)), "<f4", "<f4"],
                "offsets": [0, 12, 24, 28], "itemsize": 32})
HIT = np.dtype({"names": ["primIdx", "t", "uv"], "formats": ["<u4", "<f4", ("<f4", (2,))], "offsets": [0, 4, 8], "itemsize": 32})
TRANSFORM = np.dtype({"names": ["translation", "scale", "quat"], "formats": [("<f4", (3,)), ("<f4", (3,)), ("<f4", (4,))],
                      "offsets": [0, 16, 32], "itemsize": 64})
CAMERA = np.dtype({"names": ["eye", "quat", "fov", "near", "far"], "formats": [("<f4", (4,)), ("<f4", (4,)), "<f4", "<f4", "<f4"],
                   "offsets": [0, 16, 32, 36, 40], "itemsize": 64})

CLUSTER = np.dtype({"names": ["lo", "hi", "node", "mn", "mx"], "formats": ["<u4", "<u4", "<u4", ("<f4", (3,)), ("<f4", (3,))],
                    "offsets": [0, 4, 8, 16, 28], "itemsize": 48})

assert TRIANGLE.itemsize == 64 and AABB.itemsize == 24 and BVH2_NODE.itemsize == 32 and BVH4_NODE.itemsize == 128
assert PRIM_REF.itemsize == 28 and PRIM_NODE.itemsize == 8 and RAY.itemsize == 32 and HIT.itemsize == 32


def triangles_from_array(v):
    """(n,3,3) or (n,9) float32 -> TRIANGLE[n] (padding zeroed)."""
    v = np.asarray(v, dtype=np.float32).reshape(-1, 3, 3)
    out = np.zeros(v.shape[0], dtype=TRIANGLE)
    out["v"] = v
    return out


def make_transform(translation, scale, quat):
    t = np.zeros(1, dtype=TRANSFORM)
    t["translation"] = translation; t["scale"] = scale; t["quat"] = quat
    return t


def make_camera(eye, quat, fov, near=0.0, far=100000.0):
    c = np.zeros(1, dtype=CAMERA)
    c["eye"] = eye; c["quat"] = quat; c["fov"] = fov; c["near"] = near; c["far"] = far
    return c
