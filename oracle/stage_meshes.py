"""Stage the reference's test meshes as raw triangle arrays under oracle/_ref/meshes/ (git-ignored, but shipped
to the GPU box by gpurun) using the reference's own OBJ loader (MeshLoader::loadScene, Utility.cpp:614-759) so the
triangle order — and therefore indices, tie order and hashes — is the reference's.  Run in the authoring
container only (needs /root/reference).  File format: float32 little-endian, N x 9 (v1 v2 v3)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref  # noqa: E402

MESHES = {"cornellbox": "cornellbox/cornellBox.obj", "bunny": "bunny/bunny.obj", "sponza": "sponza/sponza.obj", "buddha": "buddha/buddha.obj"}


def main(names):
    out = os.path.join(HERE, "_ref", "meshes")
    os.makedirs(out, exist_ok=True)
    for name in names:
        path = os.path.join("/root/reference/src/Meshes", MESHES[name])
        tris = ref.load_obj(path, os.path.dirname(path) + "/")
        arr = np.ascontiguousarray(tris["v"].reshape(-1, 9))
        arr.tofile(os.path.join(out, name + ".tri"))
        print(name, arr.shape[0], "triangles")


if __name__ == "__main__":
    main(sys.argv[1:] or ["cornellbox", "bunny", "sponza"])
