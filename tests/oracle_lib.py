"""ctypes loader for the CPU oracle (oracle/libvxp_oracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libvxp_oracle.so")
MAX_QUADS = 98304
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    lib = C.CDLL(SO)
    lib.vxo_mix32.restype = C.c_uint32
    lib.vxo_mix32.argtypes = [C.c_uint32]
    lib.vxo_height.restype = C.c_int32
    lib.vxo_height.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    lib.vxo_block.restype = C.c_uint16
    lib.vxo_block.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32]
    lib.vxo_terrain_fill.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.vxo_terrain_fill_batch.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    lib.vxo_pattern_fill.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.vxo_mesh_chunk_keys.restype = C.c_uint32
    lib.vxo_mesh_chunk_keys.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.vxo_sort_keys.argtypes = [C.c_void_p, C.c_uint32]
    lib.vxo_hash_keys.restype = C.c_uint64
    lib.vxo_hash_keys.argtypes = [C.c_void_p, C.c_uint32]
    lib.vxo_expand_quad.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.vxo_mesh_batch.restype = C.c_int
    lib.vxo_mesh_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.vxo_mesh_batch_hash.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p]
    lib.vxo_max_threads.restype = C.c_int
    lib.vxo_terrain_world_hash.restype = C.c_int
    lib.vxo_terrain_world_hash.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def max_threads() -> int:
    return int(load().vxo_max_threads())


def height(seed, X, Z) -> int:
    return int(load().vxo_height(seed, X, Z))


def terrain_fill(seed, cx, cy, cz) -> np.ndarray:
    out = np.empty(32768, dtype=np.uint16)
    load().vxo_terrain_fill(seed, cx, cy, cz, out.ctypes.data)
    return out


def terrain_fill_batch(seed, coords: np.ndarray, threads: int = 0) -> np.ndarray:
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    out = np.empty((coords.shape[0], 32768), dtype=np.uint16)
    load().vxo_terrain_fill_batch(seed, coords.ctypes.data, coords.shape[0], threads or max_threads(),
                                  out.ctypes.data)
    return out


def pattern_fill(pattern: int, cx, cy, cz) -> np.ndarray:
    out = np.empty(32768, dtype=np.uint16)
    load().vxo_pattern_fill(pattern, cx, cy, cz, out.ctypes.data)
    return out


def mesh_chunk_keys(vox: np.ndarray, nbs, mode: int) -> np.ndarray:
    """Sorted canonical keys of one chunk; nbs = list of 6 arrays/None."""
    vox = np.ascontiguousarray(vox, dtype=np.uint16).reshape(-1)
    keep = [None if a is None else np.ascontiguousarray(a, dtype=np.uint16).reshape(-1) for a in (nbs or [None] * 6)]
    ptrs = (C.c_void_p * 6)(*[None if a is None else a.ctypes.data for a in keep])
    keys = np.empty(MAX_QUADS, dtype=np.uint64)
    n = load().vxo_mesh_chunk_keys(vox.ctypes.data, ptrs, mode, keys.ctypes.data)
    keys = keys[:n].copy()
    keys.sort()
    return keys


def expand_quad(key: int, q: int):
    v = np.empty(4, dtype=np.uint64)
    i = np.empty(6, dtype=np.uint32)
    load().vxo_expand_quad(int(key), q, v.ctypes.data, i.ctypes.data)
    return v, i


def mesh_batch(vox: np.ndarray, nbr6, mode: int, threads: int = 0, capacity: int | None = None):
    """Returns (verts u64[4Q], indices u32[6Q], meta u32[n,2], total, overflow)."""
    vox = np.ascontiguousarray(vox, dtype=np.uint16).reshape(-1, 32768)
    n = vox.shape[0]
    nptr = None
    if nbr6 is not None:
        nbr6 = np.ascontiguousarray(nbr6, dtype=np.int32).reshape(n, 6)
        nptr = nbr6.ctypes.data
    if capacity is None:
        hashes, counts = mesh_batch_hash(vox, nbr6, mode, threads)
        capacity = int(counts.sum())
    verts = np.empty(4 * max(capacity, 1), dtype=np.uint64)
    indices = np.empty(6 * max(capacity, 1), dtype=np.uint32)
    meta = np.zeros((n, 2), dtype=np.uint32)
    total = C.c_uint64(0)
    ovf = load().vxo_mesh_batch(vox.ctypes.data, nptr, n, mode, threads or max_threads(), verts.ctypes.data,
                                indices.ctypes.data, capacity, meta.ctypes.data, C.byref(total))
    return verts, indices, meta, int(total.value), bool(ovf)


def mesh_batch_hash(vox: np.ndarray, nbr6, mode: int, threads: int = 0):
    vox = np.ascontiguousarray(vox, dtype=np.uint16).reshape(-1, 32768)
    n = vox.shape[0]
    nptr = None
    if nbr6 is not None:
        nbr6 = np.ascontiguousarray(nbr6, dtype=np.int32).reshape(n, 6)
        nptr = nbr6.ctypes.data
    hashes = np.empty(n, dtype=np.uint64)
    counts = np.empty(n, dtype=np.uint32)
    load().vxo_mesh_batch_hash(vox.ctypes.data, nptr, n, mode, threads or max_threads(), hashes.ctypes.data,
                               counts.ctypes.data)
    return hashes, counts


def terrain_world_hash(seed, lo, hi, mode: int, threads: int = 0):
    """(hashes u64, counts u32) of the chunks of the box [lo, hi) of the infinite terrain, x-fastest order."""
    lo = np.ascontiguousarray(lo, dtype=np.int32)
    hi = np.ascontiguousarray(hi, dtype=np.int32)
    n = int(np.prod(np.maximum(hi - lo, 0)))
    hashes = np.empty(n, dtype=np.uint64)
    counts = np.empty(n, dtype=np.uint32)
    rc = load().vxo_terrain_world_hash(seed & (2**64 - 1), lo.ctypes.data, hi.ctypes.data, mode, threads or max_threads(),
                                       hashes.ctypes.data, counts.ctypes.data)
    if rc != 0:
        raise MemoryError("vxo_terrain_world_hash")
    return hashes, counts
