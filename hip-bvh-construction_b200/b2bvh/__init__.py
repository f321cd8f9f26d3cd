"""b2bvh — Python mirror of the reference's builder objects (TwoPassLbvh / SinglePassLbvh / PLOCNew / HPLOC) on top of
the C ABI of libb2bvh.so (include/b2bvh.h).  The library is hand-written sm_100a CUDA; there is no CPU fallback: importing
`b2bvh.capi` without the built library, or creating a Context without a B200, raises."""
from . import types  # noqa: F401
