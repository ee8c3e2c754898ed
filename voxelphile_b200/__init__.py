"""voxelphile_b200 — B200 (sm_100a) chunk terrain fill and meshing behind a C ABI.

Scope: SURVEY.md §8 / SPEC.md v1 only.  The compute lives in ``libvxp.so``
(``csrc/``); this package is the thin host mirror used by tests and bench.
"""
from .mesher import (CULLED, GREEDY, CHUNK_DIM, CHUNK_VOXELS, EDIT_DTYPE, DeviceMesh, HostMesh, Mesher, MeshError,
                     grid_coords, neighbour_table)
from ._lib import PATTERN_CHECKER, PATTERN_SOLID4, LIB_PATH

__all__ = ["CULLED", "GREEDY", "CHUNK_DIM", "CHUNK_VOXELS", "EDIT_DTYPE", "DeviceMesh", "HostMesh", "Mesher", "MeshError",
           "grid_coords", "neighbour_table", "PATTERN_CHECKER", "PATTERN_SOLID4", "LIB_PATH"]
