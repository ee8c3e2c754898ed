"""Pure-Python restatement of SPEC.md v1, written literally from the declarative
text (per-id greedy sets, SPEC §4) and independent of oracle/vxp_oracle.c.

Slow by design; used on small cases to cross-pin the C oracle and to generate
tests/golden/ (see tests/make_golden.py).  Parity is otherwise unpinned: the
reference snapshot has no implementation of this path.
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF

DIRS = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
GRAD = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, 1), (1, -1), (-1, -1)]


def mix32(h: int) -> int:
    h &= M32
    h ^= h >> 16
    h = (h * 0x7FEB352D) & M32
    h ^= h >> 15
    h = (h * 0x846CA68B) & M32
    h ^= h >> 16
    return h


def mix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def height(seed: int, X: int, Z: int) -> int:
    total = 0
    for o in range(4):
        S = 7 - o
        A = 192 >> o
        key = mix32(((seed >> 32) & M32) ^ mix32((seed & M32) + 0x9E3779B9 * (o + 1)))
        Xo, Zo = X + 37 * o, Z + 59 * o
        ix, iz = Xo >> S, Zo >> S  # Python >> on ints is floor
        dx = (Xo & ((1 << S) - 1)) << (8 - S)
        dz = (Zo & ((1 << S) - 1)) << (8 - S)

        def dot(i, j, ax, az):
            g = mix32(mix32(key + (i & M32) * 0x85EBCA6B) ^ (((j & M32) * 0xC2B2AE35) & M32)) & 7
            return GRAD[g][0] * ax + GRAD[g][1] * az

        d00 = dot(ix, iz, dx, dz)
        d10 = dot(ix + 1, iz, dx - 256, dz)
        d01 = dot(ix, iz + 1, dx, dz - 256)
        d11 = dot(ix + 1, iz + 1, dx - 256, dz - 256)
        fade = lambda k: (k * k * (768 - 2 * k)) >> 8
        n0 = d00 * 65536 + fade(dx) * (d10 - d00)
        n1 = d01 * 65536 + fade(dx) * (d11 - d01)
        n = n0 * 65536 + fade(dz) * (n1 - n0)
        total += n * A
    return 256 + (total >> 40)


def block(seed: int, X: int, Y: int, Z: int) -> int:
    h = height(seed, X, Z)
    if Y > h:
        return 0
    if Y == h:
        return 1
    if Y >= h - 3:
        return 2
    return 4 if ((Y >> 3) & 1) else 3


def terrain_chunk(seed: int, cx: int, cy: int, cz: int) -> np.ndarray:
    out = np.zeros((32, 32, 32), dtype=np.uint16)  # [z][y][x]
    for z in range(32):
        for x in range(32):
            h = height(seed, 32 * cx + x, 32 * cz + z)
            for y in range(32):
                Y = 32 * cy + y
                if Y > h:
                    b = 0
                elif Y == h:
                    b = 1
                elif Y >= h - 3:
                    b = 2
                else:
                    b = 4 if ((Y >> 3) & 1) else 3
                out[z, y, x] = b
    return out.reshape(-1)


def pattern_chunk(pattern: str, cx: int, cy: int, cz: int) -> np.ndarray:
    out = np.zeros((32, 32, 32), dtype=np.uint16)
    for z in range(32):
        for y in range(32):
            for x in range(32):
                X, Y, Z = 32 * cx + x, 32 * cy + y, 32 * cz + z
                if pattern == "checker":
                    out[z, y, x] = (X + Y + Z) & 1
                else:
                    out[z, y, x] = 1 + (mix32(mix32(mix32(X & M32) ^ (Y & M32)) ^ (Z & M32)) & 3)
    return out.reshape(-1)


def _voxel(vox, nbs, x, y, z):
    """vox: flat u16[32768]; nbs: list of 6 flat arrays or None."""
    src = vox
    if x < 0:
        src, x = nbs[0], x + 32
    elif x > 31:
        src, x = nbs[1], x - 32
    elif y < 0:
        src, y = nbs[2], y + 32
    elif y > 31:
        src, y = nbs[3], y - 32
    elif z < 0:
        src, z = nbs[4], z + 32
    elif z > 31:
        src, z = nbs[5], z - 32
    if src is None:
        return 0
    return int(src[x + 32 * y + 1024 * z])


def _xyz(axis, s, u, v):
    if axis == 0:
        return s, u, v
    if axis == 1:
        return u, s, v
    return u, v, s


def make_key(f, s, b, u, v, w, h) -> int:
    return u | v << 6 | s << 12 | f << 18 | (w - 1) << 21 | (h - 1) << 26 | b << 31


def mesh_chunk_keys(vox, nbs, greedy: bool) -> list[int]:
    """Sorted canonical keys per SPEC §4, following the per-id declarative rule."""
    nbs = nbs or [None] * 6
    keys = []
    for f in range(6):
        axis = f >> 1
        d = DIRS[f]
        for s in range(32):
            per_id: dict[int, set] = {}
            for v in range(32):
                for u in range(32):
                    x, y, z = _xyz(axis, s, u, v)
                    b = int(vox[x + 32 * y + 1024 * z])
                    if b and _voxel(vox, nbs, x + d[0], y + d[1], z + d[2]) == 0:
                        per_id.setdefault(b, set()).add((v, u))
            for b, M in per_id.items():
                if not greedy:
                    keys.extend(make_key(f, s, b, u, v, 1, 1) for (v, u) in M)
                    continue
                while M:
                    v, u = min(M)
                    w = 1
                    while (v, u + w) in M:
                        w += 1
                    h = 1
                    while all((v + h, u + k) in M for k in range(w)):
                        h += 1
                    for j in range(h):
                        for k in range(w):
                            M.remove((v + j, u + k))
                    keys.append(make_key(f, s, b, u, v, w, h))
    keys.sort()
    return keys


def hash_keys(keys) -> int:
    acc = 0
    for k in keys:
        acc = (acc + mix64(int(k))) & M64
    return acc


def expand_quad(key: int, q: int):
    """SPEC §5 -> (4 vertices as u64, 6 indices)."""
    u, v, s = key & 63, (key >> 6) & 63, (key >> 12) & 63
    f, w, h = (key >> 18) & 7, ((key >> 21) & 31) + 1, ((key >> 26) & 31) + 1
    b = (key >> 31) & 0xFFFF
    axis, along = f >> 1, s + (f & 1)
    corners = [(u, v), (u + w, v), (u + w, v + h), (u, v + h)]
    order = [0, 3, 2, 1] if f in (0, 3, 4) else [0, 1, 2, 3]
    verts = []
    for k in range(4):
        cu, cv = corners[order[k]]
        x, y, z = _xyz(axis, along, cu, cv)
        w0 = x | y << 6 | z << 12 | f << 18 | k << 21
        w1 = b | (w - 1) << 16 | (h - 1) << 21
        verts.append(w0 | w1 << 32)
    idx = [4 * q + t for t in (0, 1, 2, 2, 3, 0)]
    return verts, idx


# ----------------------------------------------------------------- numpy helpers
# Vectorised decoders used by the GPU parity tests (not part of the restatement).

def expand_keys(keys: np.ndarray) -> np.ndarray:
    """SPEC §5 for an array of canonical keys at once: (Q,4) u64 vertices (vectorised expand_quad)."""
    k = np.asarray(keys, dtype=np.uint64).astype(np.int64)
    u, v, s = k & 63, (k >> 6) & 63, (k >> 12) & 63
    f, w, h = (k >> 18) & 7, ((k >> 21) & 31) + 1, ((k >> 26) & 31) + 1
    b = (k >> 31) & 0xFFFF
    axis, along = f >> 1, s + (f & 1)
    flip = (f == 0) | (f == 3) | (f == 4)
    cu = np.stack([u, u + w, u + w, u], axis=1)          # corners c0..c3
    cv = np.stack([v, v, v + h, v + h], axis=1)
    order = np.where(flip[:, None], np.array([0, 3, 2, 1]), np.array([0, 1, 2, 3]))
    cu, cv = np.take_along_axis(cu, order, 1), np.take_along_axis(cv, order, 1)
    a, al = axis[:, None], along[:, None]
    x = np.where(a == 0, al, cu)
    y = np.where(a == 0, cu, np.where(a == 1, al, cv))
    z = np.where(a == 2, al, cv)
    w0 = x | y << 6 | z << 12 | f[:, None] << 18 | np.arange(4)[None, :] << 21
    w1 = (b | (w - 1) << 16 | (h - 1) << 21)[:, None]
    return (w0 | w1 << 32).astype(np.uint64)


def keys_from_verts(verts: np.ndarray) -> np.ndarray:
    """Recover canonical keys from a (Q,4) u64 vertex array (SPEC §5 inverse)."""
    verts = np.asarray(verts, dtype=np.uint64).reshape(-1, 4)
    v0 = verts[:, 0]
    w0 = (v0 & np.uint64(M32)).astype(np.int64)
    w1 = (v0 >> np.uint64(32)).astype(np.int64)
    x, y, z = w0 & 63, (w0 >> 6) & 63, (w0 >> 12) & 63
    f = (w0 >> 18) & 7
    b = w1 & 0xFFFF
    wm1, hm1 = (w1 >> 16) & 31, (w1 >> 21) & 31
    axis = f >> 1
    pos = np.stack([x, y, z], axis=1)
    rows = np.arange(len(f))
    along = pos[rows, axis] - (f & 1)
    ua = np.where(axis == 0, 1, 0)
    va = np.where(axis == 2, 1, 2)
    u, v = pos[rows, ua], pos[rows, va]  # vertex 0 is always corner c0
    key = u | v << 6 | along << 12 | f << 18 | wm1 << 21 | hm1 << 26 | b << 31
    return key.astype(np.uint64)


# --------------------------------------------------------------------------- SPEC §9 edits
def apply_edits(vox: np.ndarray, nbr, edits, n_mesh=None) -> list[int]:
    """Apply edits (iterable of (chunk, x, y, z, id)) to vox [n_stored, 32768] in place; return the sorted
    list of invalidated chunks (SPEC §9)."""
    n = vox.shape[0] if n_mesh is None else n_mesh
    dirty = set()
    for c, x, y, z, b in edits:
        vox[c, x + 32 * y + 1024 * z] = b
        dirty.add(int(c))
        if nbr is None:
            continue
        for coord, lo_f in ((x, 0), (y, 2), (z, 4)):
            for f in ((lo_f,) if coord == 0 else (lo_f + 1,) if coord == 31 else ()):
                nb = int(nbr[c][f])
                if 0 <= nb < n:
                    dirty.add(nb)
    return sorted(dirty)


# --------------------------------------------------------------------------- SPEC §10 wire format
PACK_MAGIC = 0x31505856


def pack_chunk(vox: np.ndarray) -> bytes:
    """Canonical record of one chunk (SPEC §10)."""
    vox = np.ascontiguousarray(vox, dtype=np.uint16).reshape(-1)
    pal = np.unique(vox)
    k = len(pal)
    bits = 16 if k > 16 else 0 if k <= 1 else 1 if k <= 2 else 2 if k <= 4 else 4
    head = np.zeros(4, dtype=np.uint32)
    head[0], head[1], head[2] = PACK_MAGIC, bits | ((0 if bits == 16 else k) << 16), 4096 * bits
    palette = np.zeros(16, dtype=np.uint16)
    if bits != 16:
        palette[:k] = pal
    if bits == 0:
        data = b""
    elif bits == 16:
        data = vox.tobytes()
    else:
        idx = np.searchsorted(pal, vox).astype(np.uint64)
        per = 64 // bits                                    # voxels per u64 word
        shifts = (np.arange(per, dtype=np.uint64) * np.uint64(bits))
        data = (idx.reshape(-1, per) << shifts).sum(axis=1, dtype=np.uint64).astype("<u8").tobytes()
    return head.astype("<u4").tobytes() + palette.astype("<u2").tobytes() + data


def unpack_record(rec: bytes) -> np.ndarray:
    head = np.frombuffer(rec[:16], dtype="<u4")
    bits, k = int(head[1]) & 0xFFFF, int(head[1]) >> 16
    assert head[0] == PACK_MAGIC and bits in (0, 1, 2, 4, 16) and head[2] == 4096 * bits
    palette = np.frombuffer(rec[16:48], dtype="<u2")
    if bits == 16:
        assert k == 0
        return np.frombuffer(rec[48:48 + 65536], dtype="<u2").copy()
    assert 1 <= k <= 16 and k <= (1 << bits)
    if bits == 0:
        return np.full(32768, palette[0], dtype=np.uint16)
    words = np.frombuffer(rec[48:48 + 4096 * bits], dtype="<u8")
    per = 64 // bits
    shifts = (np.arange(per, dtype=np.uint64) * np.uint64(bits))
    idx = ((words[:, None] >> shifts) & np.uint64((1 << bits) - 1)).reshape(-1)
    return palette[idx.astype(np.int64)]


def record_size(rec_head: bytes) -> int:
    return 48 + int(np.frombuffer(rec_head[8:12], dtype="<u4")[0])


# --------------------------------------------------------------------------- SPEC §11 level of detail
def downsample(children) -> np.ndarray:
    """children: list of 8 arrays [32768] or None, index dx + 2 dy + 4 dz -> one chunk [32768]."""
    region = np.zeros((64, 64, 64), dtype=np.uint16)              # [z][y][x]
    for c, v in enumerate(children):
        if v is None:
            continue
        dx, dy, dz = c & 1, (c >> 1) & 1, c >> 2
        region[32 * dz:32 * dz + 32, 32 * dy:32 * dy + 32, 32 * dx:32 * dx + 32] = np.asarray(v, np.uint16).reshape(32, 32, 32)
    out = np.zeros((32, 32, 32), dtype=np.uint16)
    for j, k, i in ((1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)):   # (y, z, x) offsets
        cand = region[k::2, j::2, i::2]
        out = np.where(out == 0, cand, out)
    return out.reshape(-1)
