"""Host-side chunk / mesher surface over the C ABI (include/vxp.h).

The reference snapshot defines no chunk or mesher API (its only public items are
``xenotech::main``, the login enums and ``User`` — src/lib.rs:9,31-47), so this
module defines the surface the Rust crate in ``rust/vxp`` mirrors: chunk
coordinate -> voxel array -> mesh buffers.  Errors follow the reference's
payload-free enum style (src/net/udp/mod.rs:16-28): ``MeshError.kind`` is one of
a fixed set of names and carries no payload beyond the library's detail string.

PyTorch is used only to own device memory and streams; all compute is in
libvxp.so (hand-written sm_100a kernels).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

CULLED, GREEDY = _lib.MODE_CULLED, _lib.MODE_GREEDY
CHUNK_DIM = 32
CHUNK_VOXELS = _lib.CHUNK_VOXELS
# vxp_edit (include/vxp.h): voxel (x,y,z) of stored chunk `chunk` becomes `id`
EDIT_DTYPE = np.dtype([("chunk", np.uint32), ("x", np.uint16), ("y", np.uint16), ("z", np.uint16), ("id", np.uint16)])

_KINDS = {
    _lib.VXP_ERR_BAD_ARG: "BadArg",
    _lib.VXP_ERR_NO_DEVICE: "NoDevice",
    _lib.VXP_ERR_OOM: "OutOfMemory",
    _lib.VXP_ERR_CAPACITY: "Capacity",
    _lib.VXP_ERR_CUDA: "Cuda",
}


class MeshError(RuntimeError):
    """Payload-free error kind (``BadArg | NoDevice | OutOfMemory | Capacity | Cuda``)."""

    def __init__(self, status: int, detail: str = ""):
        self.status = status
        self.kind = _KINDS.get(status, "Unknown")
        super().__init__(f"{self.kind}: {detail}" if detail else self.kind)


def neighbour_table(coords: np.ndarray) -> np.ndarray:
    """``nbr[i][f]`` = index of the chunk at ``coords[i] + dir_f`` or -1 (SPEC §1). Host-only."""
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
    nbr = np.empty((coords.shape[0], 6), dtype=np.int32)
    st = _lib.load().vxp_nbr6_from_coords(coords.ctypes.data, coords.shape[0], nbr.ctypes.data)
    if st != _lib.VXP_OK:
        raise MeshError(st, "neighbour_table")
    return nbr


def grid_coords(lo, hi) -> np.ndarray:
    """Chunk coordinates of the box [lo, hi) in x-fastest order, as int32 [n,3]."""
    lo, hi = np.asarray(lo), np.asarray(hi)
    z, y, x = np.meshgrid(np.arange(lo[2], hi[2]), np.arange(lo[1], hi[1]), np.arange(lo[0], hi[0]), indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int32)


@dataclass
class DeviceMesh:
    """Batch mesh in device memory (SPEC §5). ``verts`` holds u64 bit patterns in int64."""

    verts: "torch.Tensor"    # int64 [4*capacity]
    indices: "Optional[torch.Tensor]"  # int32 [6*capacity], or None (not written)
    meta: "torch.Tensor"     # int32 [n,2]  (first_quad, quad_count)
    total: "torch.Tensor"    # int64 [1]
    capacity: int

    def to_host(self) -> "HostMesh":
        total = int(self.total.item())
        if total > self.capacity:
            raise MeshError(_lib.VXP_ERR_CAPACITY, f"need {total} quads, capacity {self.capacity}")
        return HostMesh(
            verts=self.verts[: 4 * total].cpu().numpy().view(np.uint64),
            indices=self.indices[: 6 * total].cpu().numpy().view(np.uint32) if self.indices is not None else np.empty(0, np.uint32),
            meta=self.meta.cpu().numpy().view(np.uint32).reshape(-1, 2),
            total_quads=total,
        )


@dataclass
class HostMesh:
    verts: np.ndarray    # uint64 [4*Q]
    indices: np.ndarray  # uint32 [6*Q]
    meta: np.ndarray     # uint32 [n,2]
    total_quads: int

    def chunk(self, i: int):
        """(verts[cnt,4], indices[cnt,6]) of chunk i."""
        first, cnt = int(self.meta[i, 0]), int(self.meta[i, 1])
        ix = self.indices[6 * first : 6 * (first + cnt)].reshape(cnt, 6) if self.indices.size else self.indices.reshape(0, 6)
        return self.verts[4 * first : 4 * (first + cnt)].reshape(cnt, 4), ix


class Mesher:
    """One context = one GPU.  Mirrors ``vxp_ctx`` 1:1.

    Streams.  The device-level methods take and return torch tensors, so by default every call runs on torch's
    CURRENT stream of the device (looked up per call): allocations, fills, kernels and read-backs are then ordered
    the way torch orders its own work, and the caching allocator never recycles a block a kernel still uses.
    ``use_stream(s)`` pins one stream instead; ``use_stream(None)`` selects the context's private non-blocking
    stream, which has no ordering with torch's streams - only for callers that pass raw host pointers (the
    host-level methods synchronise themselves) or that order device tensors with events of their own."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        st = self._lib.vxp_ctx_create(device, C.byref(h))
        if st != _lib.VXP_OK:
            raise MeshError(st, self._lib.vxp_status_str(st).decode())
        self._h = h
        self._follow = True          # run on torch's current stream, re-read at every device-level call
        self._stream_ptr = None      # what the context was last told (None: its own stream)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vxp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != _lib.VXP_OK:
            detail = self._lib.vxp_last_error(self._h).decode() if st == _lib.VXP_ERR_CUDA else \
                self._lib.vxp_status_str(st).decode()
            raise MeshError(st, detail)

    @property
    def launches(self) -> int:
        return int(self._lib.vxp_launch_count(self._h))

    def set_overlap(self, enabled: bool):
        """vxp_ctx_set_overlap: consecutive mesh() calls may overlap on the device; the caller alternates
        between two output meshes and does not rewrite a call's inputs behind the previous call."""
        self._check(self._lib.vxp_ctx_set_overlap(self._h, 1 if enabled else 0))

    def set_host_indices(self, enabled: bool):
        """vxp_ctx_set_host_indices: with False the host-level calls neither produce nor download the index buffer
        (``HostMesh.indices`` is then empty: SPEC §5 makes it 4q + {0,1,2,2,3,0})."""
        self._check(self._lib.vxp_ctx_set_host_indices(self._h, 1 if enabled else 0))

    def use_stream(self, stream: Optional["torch.cuda.Stream"]):
        """Pin a torch stream; ``None`` selects the context's own non-blocking stream (see the class docstring)."""
        self._follow = False
        if stream is None:
            self._check(self._lib.vxp_ctx_reset_stream(self._h))
            self._stream_ptr = None
        else:
            self._check(self._lib.vxp_ctx_set_stream(self._h, stream.cuda_stream or None))
            self._stream_ptr = stream.cuda_stream

    def use_current_torch_stream(self):
        """Back to the default: every device-level call runs on torch's current stream at the time of the call."""
        self._follow = True
        self._on_torch_stream()

    def _on_torch_stream(self):
        if not self._follow:
            return
        import torch
        ptr = torch.cuda.current_stream(self.device).cuda_stream
        if ptr != self._stream_ptr:
            self._check(self._lib.vxp_ctx_set_stream(self._h, ptr or None))
            self._stream_ptr = ptr

    def synchronize(self):
        self._check(self._lib.vxp_ctx_synchronize(self._h))

    # ------------------------------------------------------------ device level (torch tensors)
    def _dev(self):
        import torch
        return torch.device("cuda", self.device)

    def alloc_mesh(self, n: int, capacity: int, indices: bool = True) -> DeviceMesh:
        """indices=False: no index buffer (vxp_mesh_dev.indices == NULL: the kernels do not write indices)."""
        self._on_torch_stream()
        import torch
        d = self._dev()
        return DeviceMesh(
            verts=torch.empty(4 * max(capacity, 1), dtype=torch.int64, device=d),
            indices=torch.empty(6 * max(capacity, 1), dtype=torch.int32, device=d) if indices else None,
            meta=torch.zeros((n, 2), dtype=torch.int32, device=d),
            total=torch.zeros(1, dtype=torch.int64, device=d),
            capacity=capacity,
        )

    @staticmethod
    def _mesh_struct(m: DeviceMesh) -> _lib.MeshDev:
        return _lib.MeshDev(m.verts.data_ptr(), m.indices.data_ptr() if m.indices is not None else None, m.meta.data_ptr(),
                            m.total.data_ptr(), m.capacity)

    def terrain_fill(self, coords: "torch.Tensor", seed: int, out: Optional["torch.Tensor"] = None):
        """coords: int32 [n,3] on this device -> int16 [n,32768] (u16 bit patterns)."""
        self._on_torch_stream()
        import torch
        n = coords.shape[0]
        assert coords.dtype == torch.int32 and coords.is_contiguous() and coords.is_cuda
        if out is None:
            out = torch.empty((n, CHUNK_VOXELS), dtype=torch.int16, device=coords.device)
        self._check(self._lib.vxp_terrain_fill_batch(self._h, coords.data_ptr(), n, seed & (2**64 - 1),
                                                     out.data_ptr()))
        return out

    def pattern_fill(self, coords: "torch.Tensor", pattern: int, out: Optional["torch.Tensor"] = None):
        self._on_torch_stream()
        import torch
        n = coords.shape[0]
        assert coords.dtype == torch.int32 and coords.is_contiguous() and coords.is_cuda
        if out is None:
            out = torch.empty((n, CHUNK_VOXELS), dtype=torch.int16, device=coords.device)
        self._check(self._lib.vxp_pattern_fill_batch(self._h, coords.data_ptr(), n, pattern, out.data_ptr()))
        return out

    def mesh(self, voxels: "torch.Tensor", nbr6: Optional["torch.Tensor"], mode: int,
             out: Optional[DeviceMesh] = None, capacity: Optional[int] = None,
             n_mesh: Optional[int] = None) -> DeviceMesh:
        """voxels: 16-bit [n_stored,32768] device tensor; nbr6: int32 [n,6] or None. Asynchronous.
        Only the first ``n_mesh`` (default all) chunks are meshed; the rest are read-only halo chunks."""
        self._on_torch_stream()
        n = voxels.shape[0] if n_mesh is None else n_mesh
        assert voxels.is_cuda and voxels.is_contiguous() and voxels.element_size() == 2
        if nbr6 is not None:
            assert nbr6.is_cuda and nbr6.is_contiguous() and nbr6.shape == (n, 6)
        if out is None:
            out = self.alloc_mesh(n, capacity if capacity is not None else n * 1024)
        s = self._mesh_struct(out)
        self._check(self._lib.vxp_mesh_batch(self._h, voxels.data_ptr(),
                                             nbr6.data_ptr() if nbr6 is not None else None, n, mode, C.byref(s)))
        return out

    def terrain_mesh(self, coords: "torch.Tensor", seed: int, mode: int, out: Optional[DeviceMesh] = None,
                     capacity: Optional[int] = None) -> DeviceMesh:
        self._on_torch_stream()
        n = coords.shape[0]
        if out is None:
            out = self.alloc_mesh(n, capacity if capacity is not None else n * 256)
        s = self._mesh_struct(out)
        self._check(self._lib.vxp_terrain_mesh_fused(self._h, coords.data_ptr(), n, seed & (2**64 - 1), mode,
                                                     C.byref(s)))
        return out

    # ------------------------------------------------------------ incremental re-mesh
    def mesh_list(self, voxels: "torch.Tensor", nbr6: Optional["torch.Tensor"], chunk_list: "torch.Tensor", mode: int,
                  out: Optional[DeviceMesh] = None, capacity: Optional[int] = None) -> DeviceMesh:
        """vxp_mesh_list: mesh the chunks named by ``chunk_list`` (int32/uint32 device tensor); meta entry k
        belongs to chunk chunk_list[k]."""
        self._on_torch_stream()
        n, count = voxels.shape[0], int(chunk_list.shape[0])
        assert chunk_list.is_cuda and chunk_list.is_contiguous() and chunk_list.element_size() == 4
        if out is None:
            out = self.alloc_mesh(count, capacity if capacity is not None else count * 4096)
        s = self._mesh_struct(out)
        self._check(self._lib.vxp_mesh_list(self._h, voxels.data_ptr(), nbr6.data_ptr() if nbr6 is not None else None,
                                            n, chunk_list.data_ptr(), count, mode, C.byref(s)))
        return out

    def remesh_edits(self, voxels: "torch.Tensor", nbr6: Optional["torch.Tensor"], edits: "torch.Tensor", mode: int,
                     capacity: Optional[int] = None):
        """vxp_remesh_edits: apply ``edits`` (device tensor of EDIT_DTYPE records viewed as int32 [m,3]) to
        ``voxels`` in place and re-mesh the invalidated chunks.  Returns (dirty list tensor, count tensor, DeviceMesh);
        nothing is synchronised."""
        self._on_torch_stream()
        import torch
        n, m = voxels.shape[0], int(edits.shape[0])
        cap = min(n, 7 * m)
        dirty = torch.empty(max(cap, 1), dtype=torch.int32, device=voxels.device)
        count = torch.zeros(1, dtype=torch.int32, device=voxels.device)
        out = self.alloc_mesh(cap, capacity if capacity is not None else cap * 4096)
        s = self._mesh_struct(out)
        self._check(self._lib.vxp_remesh_edits(self._h, voxels.data_ptr(), nbr6.data_ptr() if nbr6 is not None else None,
                                               n, edits.data_ptr(), m, mode, dirty.data_ptr(), count.data_ptr(), C.byref(s)))
        return dirty, count, out

    def remesh_edits_host(self, edits: np.ndarray, mode: int, copy: bool = True):
        """vxp_remesh_edits_host: edit the world left resident by the last mesh_host call.  Returns
        (dirty chunk indices, HostMesh whose meta[k] belongs to dirty[k])."""
        edits = np.ascontiguousarray(edits, dtype=EDIT_DTYPE)
        h = _lib.MeshHost()
        dptr, cnt = C.c_void_p(), C.c_uint32()
        self._check(self._lib.vxp_remesh_edits_host(self._h, edits.ctypes.data, edits.shape[0], mode, C.byref(h),
                                                    C.byref(dptr), C.byref(cnt)))
        k = int(cnt.value)
        if k:
            buf = (C.c_char * (4 * k)).from_address(dptr.value)
            dirty = np.frombuffer(buf, dtype=np.uint32, count=k)
            dirty = dirty.copy() if copy else dirty
        else:
            dirty = np.empty(0, np.uint32)
        return dirty, self._host_result(h, k, copy)

    # ------------------------------------------------------------ chunk wire format (SPEC §10)
    def pack(self, voxels: "torch.Tensor", capacity_bytes: Optional[int] = None):
        """vxp_pack_chunks: (packed uint8 tensor, offsets int64 tensor [n], total-bytes int64 tensor [1]); async."""
        self._on_torch_stream()
        import torch
        n = voxels.shape[0]
        cap = n * (48 + 65536) if capacity_bytes is None else capacity_bytes
        packed = torch.empty(max(cap, 16), dtype=torch.uint8, device=voxels.device)
        offsets = torch.empty(max(n, 1), dtype=torch.int64, device=voxels.device)
        total = torch.zeros(1, dtype=torch.int64, device=voxels.device)
        self._check(self._lib.vxp_pack_chunks(self._h, voxels.data_ptr(), n, packed.data_ptr(), cap, offsets.data_ptr(),
                                              total.data_ptr()))
        return packed, offsets[:n], total

    def unpack(self, packed: "torch.Tensor", offsets: "torch.Tensor", out: Optional["torch.Tensor"] = None) -> "torch.Tensor":
        """vxp_unpack_chunks: int16 [n, 32768]; a malformed record surfaces at the next synchronize()."""
        self._on_torch_stream()
        import torch
        n = int(offsets.shape[0])
        if out is None:
            out = torch.empty((n, CHUNK_VOXELS), dtype=torch.int16, device=packed.device)
        self._check(self._lib.vxp_unpack_chunks(self._h, packed.data_ptr(), packed.numel() * packed.element_size(), offsets.data_ptr(),
                                                n, out.data_ptr()))
        return out

    def mesh_packed_host(self, packed: np.ndarray, offsets: np.ndarray, nbr6: Optional[np.ndarray], mode: int,
                         copy: bool = True, n_mesh: Optional[int] = None) -> HostMesh:
        """vxp_mesh_packed_host: packed host bytes in, host mesh out."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_stored = offsets.shape[0]
        n = n_stored if n_mesh is None else n_mesh
        nptr = None
        if nbr6 is not None:
            nbr6 = np.ascontiguousarray(nbr6, dtype=np.int32).reshape(n, 6)
            nptr = nbr6.ctypes.data
        h = _lib.MeshHost()
        self._check(self._lib.vxp_mesh_packed_host(self._h, packed.ctypes.data, packed.nbytes, offsets.ctypes.data, nptr,
                                                   n, n_stored, mode, C.byref(h)))
        return self._host_result(h, n, copy)

    def mesh_packed_host_ptr(self, packed_ptr: int, packed_bytes: int, offsets_ptr: int, nbr6_ptr: Optional[int], n: int,
                             mode: int) -> HostMesh:
        h = _lib.MeshHost()
        self._check(self._lib.vxp_mesh_packed_host(self._h, packed_ptr, packed_bytes, offsets_ptr, nbr6_ptr, n, n, mode,
                                                   C.byref(h)))
        return self._host_result(h, n, copy=False)

    # ------------------------------------------------------------ level of detail (SPEC §11)
    def downsample(self, voxels: "torch.Tensor", child8: "torch.Tensor", out: Optional["torch.Tensor"] = None) -> "torch.Tensor":
        """vxp_downsample_chunks: child8 int32 [n_out, 8] (dx + 2 dy + 4 dz, -1 = air) -> int16 [n_out, 32768]."""
        self._on_torch_stream()
        import torch
        n_out = int(child8.shape[0])
        assert child8.is_cuda and child8.is_contiguous() and child8.shape[1] == 8
        if out is None:
            out = torch.empty((n_out, CHUNK_VOXELS), dtype=torch.int16, device=voxels.device)
        self._check(self._lib.vxp_downsample_chunks(self._h, voxels.data_ptr(), child8.data_ptr(), n_out, out.data_ptr()))
        return out

    # ------------------------------------------------------------ host level (numpy in, numpy out)
    def _host_result(self, h: _lib.MeshHost, n: int, copy: bool) -> HostMesh:
        q = int(h.total_quads)

        def view(ptr, count, dtype):
            if count == 0:
                return np.empty(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dtype=dtype, count=count)
            return a.copy() if copy else a

        return HostMesh(view(h.verts, 4 * q, np.uint64), view(h.indices, 6 * q if h.indices else 0, np.uint32),
                        view(h.meta, 2 * n, np.uint32).reshape(n, 2), q)

    def mesh_host(self, voxels: np.ndarray, nbr6: Optional[np.ndarray], mode: int, copy: bool = True,
                  n_mesh: Optional[int] = None) -> HostMesh:
        """Host buffers in, host buffers out (H2D + kernel + D2H inside the call).

        ``n_mesh`` < len(voxels) marks the trailing chunks as read-only halo chunks."""
        voxels = np.ascontiguousarray(voxels, dtype=np.uint16).reshape(-1, CHUNK_VOXELS)
        n_stored = voxels.shape[0]
        n = n_stored if n_mesh is None else n_mesh
        nptr = None
        if nbr6 is not None:
            nbr6 = np.ascontiguousarray(nbr6, dtype=np.int32).reshape(n, 6)
            nptr = nbr6.ctypes.data
        h = _lib.MeshHost()
        self._check(self._lib.vxp_mesh_chunks_host(self._h, voxels.ctypes.data, nptr, n, n_stored, mode,
                                                   C.byref(h)))
        return self._host_result(h, n, copy)

    def mesh_host_ptr(self, voxels_ptr: int, nbr6_ptr: Optional[int], n: int, mode: int,
                      n_stored: Optional[int] = None) -> HostMesh:
        """Same, from raw (e.g. pinned) host pointers; the result aliases context-owned pinned memory."""
        h = _lib.MeshHost()
        self._check(self._lib.vxp_mesh_chunks_host(self._h, voxels_ptr, nbr6_ptr, n,
                                                   n if n_stored is None else n_stored, mode, C.byref(h)))
        return self._host_result(h, n, copy=False)

    def terrain_mesh_host(self, coords: np.ndarray, seed: int, mode: int, copy: bool = True) -> HostMesh:
        coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
        h = _lib.MeshHost()
        self._check(self._lib.vxp_terrain_mesh_host(self._h, coords.ctypes.data, coords.shape[0],
                                                    seed & (2**64 - 1), mode, C.byref(h)))
        return self._host_result(h, coords.shape[0], copy)

    def terrain_fill_host(self, coords: np.ndarray, seed: int) -> np.ndarray:
        coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 3)
        out = np.empty((coords.shape[0], CHUNK_VOXELS), dtype=np.uint16)
        self._check(self._lib.vxp_terrain_fill_host(self._h, coords.ctypes.data, coords.shape[0],
                                                    seed & (2**64 - 1), out.ctypes.data))
        return out
