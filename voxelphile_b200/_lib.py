"""ctypes binding of libvxp.so (include/vxp.h).  The library is built in-tree by
``voxelphile_b200/csrc/Makefile`` (``__graft_entry__.build()``); there is no
fallback of any kind: a missing library is an ImportError-grade failure.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VXP_LIB", os.path.join(_HERE, "libvxp.so"))  # VXP_LIB: profiling builds only

VXP_OK, VXP_ERR_BAD_ARG, VXP_ERR_NO_DEVICE, VXP_ERR_OOM, VXP_ERR_CAPACITY, VXP_ERR_CUDA = range(6)
MODE_CULLED, MODE_GREEDY = 0, 1
PATTERN_CHECKER, PATTERN_SOLID4 = 0, 1
CHUNK_VOXELS = 32768
MAX_QUADS_PER_CHUNK = 98304


class MeshDev(C.Structure):
    _fields_ = [
        ("verts", C.c_void_p),
        ("indices", C.c_void_p),
        ("meta", C.c_void_p),
        ("total_quads", C.c_void_p),
        ("capacity_quads", C.c_uint64),
    ]


class MeshHost(C.Structure):
    _fields_ = [
        ("verts", C.c_void_p),
        ("indices", C.c_void_p),
        ("meta", C.c_void_p),
        ("total_quads", C.c_uint64),
    ]


# name -> (restype, argtypes); kept in sync with include/vxp.h (tests/test_abi.py checks it)
SIGNATURES = {
    "vxp_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vxp_ctx_destroy": (None, [C.c_void_p]),
    "vxp_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxp_ctx_reset_stream": (C.c_int, [C.c_void_p]),
    "vxp_ctx_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "vxp_ctx_set_host_indices": (C.c_int, [C.c_void_p, C.c_int]),
    "vxp_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "vxp_status_str": (C.c_char_p, [C.c_int]),
    "vxp_last_error": (C.c_char_p, [C.c_void_p]),
    "vxp_launch_count": (C.c_uint64, [C.c_void_p]),
    "vxp_spec_version": (C.c_int, []),
    "vxp_nbr6_from_coords": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "vxp_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "vxp_host_free": (None, [C.c_void_p]),
    "vxp_terrain_fill_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p]),
    "vxp_pattern_fill_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]),
    "vxp_mesh_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(MeshDev)]),
    "vxp_terrain_mesh_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(MeshDev)]),
    "vxp_mesh_list": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int,
                      C.POINTER(MeshDev)]),
    "vxp_remesh_edits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int,
                         C.c_void_p, C.c_void_p, C.POINTER(MeshDev)]),
    "vxp_remesh_edits_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(MeshHost),
                              C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    "vxp_pack_chunks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "vxp_unpack_chunks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p]),
    "vxp_mesh_packed_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                             C.c_int, C.POINTER(MeshHost)]),
    "vxp_downsample_chunks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "vxp_terrain_fill_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p]),
    "vxp_mesh_chunks_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                             C.POINTER(MeshHost)]),
    "vxp_terrain_mesh_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(MeshHost)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libvxp.so and declare every entry point; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(voxelphile_b200 has no CPU or PyTorch fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
