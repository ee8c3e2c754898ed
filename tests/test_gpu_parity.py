"""GPU parity tests: the CUDA path, called through the C ABI (libvxp.so), against the CPU
oracle and the golden fixtures.  Bit-exact: canonical keys AND the vertex/index bytes
re-ordered by key (SPEC §5).  Comparand = in-repo oracle, SPEC v1 (parity with the
reference itself is unpinned — SURVEY.md §0)."""
import os

import numpy as np
import pytest

import oracle_lib as O
import spec_py as S
from voxelphile_b200 import EDIT_DTYPE

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "spec_v1.npz")
MODES = [(0, "culled"), (1, "greedy")]


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def mesher():
    import torch
    from voxelphile_b200 import Mesher
    assert torch.cuda.is_available()
    m = Mesher(0)
    m.use_current_torch_stream()
    yield m
    m.close()


def dev(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.view(dtype)
    return t.cuda()


def gpu_mesh(mesher, vox, nbr, mode, capacity=None, n_mesh=None):
    """Mesh on the GPU through vxp_mesh_batch; returns a HostMesh."""
    import torch
    vox = np.ascontiguousarray(vox, dtype=np.uint16).reshape(-1, 32768)
    dv = dev(vox.view(np.int16))
    dn = dev(np.asarray(nbr, dtype=np.int32)) if nbr is not None else None
    cap = capacity if capacity is not None else vox.shape[0] * 98304
    out = mesher.mesh(dv, dn, mode, capacity=cap, n_mesh=n_mesh)
    torch.cuda.synchronize()
    return out.to_host()


def check_against_oracle(host, vox, nbr, mode, expect_keys=None):
    """Every chunk: sorted keys equal, and vertex/index bytes equal the oracle's expansion."""
    n = vox.shape[0]
    assert host.meta.shape == (n, 2)
    total = 0
    used = np.zeros(host.total_quads, dtype=bool)
    for i in range(n):
        first, cnt = int(host.meta[i, 0]), int(host.meta[i, 1])
        nbs = [None if (nbr is None or nbr[i, f] < 0) else vox[nbr[i, f]] for f in range(6)]
        want = O.mesh_chunk_keys(vox[i], nbs, mode) if expect_keys is None else expect_keys[i]
        assert cnt == len(want), (i, cnt, len(want))
        total += cnt
        if cnt == 0:
            continue
        assert first + cnt <= host.total_quads
        assert not used[first:first + cnt].any(), "chunk ranges overlap"
        used[first:first + cnt] = True
        v, ix = host.chunk(i)
        got = S.keys_from_verts(v)
        order = np.argsort(got, kind="stable")
        assert np.array_equal(got[order], want), f"chunk {i}: key sets differ"
        # vertex bytes of EVERY quad, in key order, against the SPEC §5 expansion of the oracle's keys (the numpy
        # expansion is itself checked against the oracle's C expansion: tests/test_oracle.py); a sample through the
        # oracle's own vxo_expand_quad on top
        assert np.array_equal(v[order], S.expand_keys(want)), f"chunk {i}: vertex bytes"
        for q in range(0, cnt, max(1, cnt // 64)):
            assert np.array_equal(v[order[q]], O.expand_quad(int(want[q]), 0)[0]), (i, q)
        # every vertex of every quad carries the same (f, id, w, h) and corner numbers 0..3
        w0 = (v & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert np.array_equal((w0 >> 21) & 3, np.tile(np.arange(4), (cnt, 1)))
        assert (v >> np.uint64(32) == (v[:, :1] >> np.uint64(32))).all()
        assert (((w0 >> 18) & 7) == ((w0[:, :1] >> 18) & 7)).all()
        # indices are chunk-local ordinals
        want_ix = 4 * np.arange(cnt, dtype=np.uint32)[:, None] + np.array([0, 1, 2, 2, 3, 0], dtype=np.uint32)
        assert np.array_equal(ix, want_ix), f"chunk {i}: indices"
    assert total == host.total_quads and used.all()


@pytest.mark.parametrize("mode,mname", MODES)
def test_golden_cases(G, mesher, mode, mname):
    for name in [str(x) for x in G["case_names"]]:
        vox, nbr = G[f"{name}/vox"], G[f"{name}/nbr"]
        keys, offs = G[f"{name}/{mname}/keys"], G[f"{name}/{mname}/offs"]
        expect = [keys[offs[i]:offs[i + 1]] for i in range(vox.shape[0])]
        host = gpu_mesh(mesher, vox, nbr, mode)
        check_against_oracle(host, vox, nbr, mode, expect_keys=expect)


# id palettes: ids that fit the id bit planes of the greedy kernel (1..4), ids that do not (tile equality pass),
# and ids that differ ONLY above the plane bits (would merge if the planes were used for them)
PALETTES = {"small": [1, 2, 3, 4], "large": [1000, 1001, 1002, 0xFFFF], "high_bits": [3, 11, 0x103, 0x8003]}


@pytest.mark.parametrize("mode,mname", MODES)
@pytest.mark.parametrize("palette", list(PALETTES))
@pytest.mark.parametrize("p", [0.01, 0.1, 0.5, 0.9, 0.99])
def test_differential_random(mesher, mode, mname, p, palette):
    rng = np.random.default_rng(int(p * 1000) + mode)
    n = 24 if palette == "small" else 8
    ids = np.array(PALETTES[palette], dtype=np.uint16)
    vox = np.where(rng.random((n, 32768)) < p, ids[rng.integers(0, 4, (n, 32768))], 0).astype(np.uint16)
    nbr = rng.integers(-1, n, size=(n, 6)).astype(np.int32)     # arbitrary, asymmetric neighbour links
    host = gpu_mesh(mesher, vox, nbr, mode)
    check_against_oracle(host, vox, nbr, mode)


@pytest.mark.parametrize("palette", list(PALETTES))
def test_layered_ids(mesher, palette):
    """Terrain-like chunks (solid below a height field, ids in horizontal bands and vertical stripes) for every
    palette: long greedy runs that must break exactly where the id changes."""
    rng = np.random.default_rng(3)
    ids = np.array(PALETTES[palette], dtype=np.uint16)
    n = 6
    y = np.arange(32)[None, :, None]
    v = np.zeros((n, 32, 32, 32), dtype=np.uint16)              # [z][y][x]
    for i in range(n):
        h = (8 + 6 * np.sin(np.arange(32)[:, None, None] / 5.0 + i) + 6 * np.cos(np.arange(32)[None, None, :] / 7.0)).astype(int)
        band = ids[np.clip((h - y) // 2, 0, 3)] if i % 2 == 0 else ids[(np.arange(32)[None, None, :] // (2 + i)) % 4] + 0 * y
        v[i] = np.where(y <= h, band, 0)
    vox = v.reshape(n, -1)
    nbr = rng.integers(-1, n, size=(n, 6)).astype(np.int32)
    host = gpu_mesh(mesher, vox, nbr, 1)
    check_against_oracle(host, vox, nbr, 1)


def test_many_ids_and_long_runs(mesher):
    """Ids up to 0xFFFF, stripes and plateaus that stress width/height extension."""
    rng = np.random.default_rng(11)
    n = 8
    v = np.zeros((n, 32, 32, 32), dtype=np.uint16)
    v[0] = 0xFFFF
    v[1, :, :16, :] = 0x8000; v[1, :, 16:, :] = 0x7FFF
    v[2] = rng.integers(1, 3, (32, 1, 1))            # id varies along z only
    v[3] = rng.integers(1, 3, (1, 32, 1))            # along y only
    v[4] = rng.integers(1, 3, (1, 1, 32))            # along x only
    v[5, ::2] = 1                                    # alternating solid/air layers
    v[6, :, ::2] = 2
    v[7, :, :, ::2] = 3
    vox = v.reshape(n, -1)
    for mode, _ in MODES:
        host = gpu_mesh(mesher, vox, None, mode)
        check_against_oracle(host, vox, None, mode)


def test_terrain_fill_bit_exact(G, mesher):
    import torch
    from voxelphile_b200 import grid_coords
    seed = int(G["terrain_seed"][0])
    got = mesher.terrain_fill(dev(G["terrain_coords"]), seed)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy().view(np.uint16), G["terrain_vox"])
    for seed, lo in ((42, (0, 6, 0)), (1, (-3, 5, -2)), (2**63 + 12345, (100, 7, -100))):
        coords = grid_coords(lo, (lo[0] + 3, lo[1] + 5, lo[2] + 3))
        got = mesher.terrain_fill(dev(coords), seed)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy().view(np.uint16), O.terrain_fill_batch(seed, coords))


def test_pattern_fill_bit_exact(mesher):
    import torch
    from voxelphile_b200 import grid_coords
    coords = grid_coords((-1, -1, -1), (1, 1, 1))
    for pid in (0, 1):
        got = mesher.pattern_fill(dev(coords), pid)
        torch.cuda.synchronize()
        want = np.stack([O.pattern_fill(pid, *map(int, c)) for c in coords])
        assert np.array_equal(got.cpu().numpy().view(np.uint16), want)


@pytest.mark.parametrize("mode,mname", MODES)
def test_terrain_block_stored_path(mesher, mode, mname):
    """Config 1/2/3 in miniature: a block of seed-42 terrain chunks with real neighbours."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table
    coords = grid_coords((0, 5, 0), (4, 11, 4))
    nbr = neighbour_table(coords)
    dvox = mesher.terrain_fill(dev(coords), 42)
    out = mesher.mesh(dvox, dev(nbr), mode, capacity=coords.shape[0] * 4096)
    torch.cuda.synchronize()
    vox = dvox.cpu().numpy().view(np.uint16)
    check_against_oracle(out.to_host(), vox, nbr, mode)


@pytest.mark.parametrize("mode,mname", MODES)
def test_fused_terrain_mesh(mesher, mode, mname):
    """Fused path: neighbours are the infinite terrain itself."""
    import torch
    from voxelphile_b200 import grid_coords
    seed = 42
    coords = np.concatenate([grid_coords((0, 6, 0), (3, 10, 2)), np.array([[-7, 8, 3], [5, 0, 5], [5, 30, 5]], np.int32)])
    out = mesher.terrain_mesh(dev(coords), seed, mode, capacity=coords.shape[0] * 8192)
    torch.cuda.synchronize()
    host = out.to_host()
    expect = []
    for c in coords:
        c = tuple(map(int, c))
        vox = O.terrain_fill(seed, *c)
        nbs = [O.terrain_fill(seed, c[0] + d[0], c[1] + d[1], c[2] + d[2]) for d in S.DIRS]
        expect.append((vox, nbs))
    for i, (vox, nbs) in enumerate(expect):
        first, cnt = int(host.meta[i, 0]), int(host.meta[i, 1])
        want = O.mesh_chunk_keys(vox, nbs, mode)
        assert cnt == len(want), (i, cnt, len(want))
        v, ix = host.chunk(i)
        assert np.array_equal(np.sort(S.keys_from_verts(v)), want), i
    assert host.total_quads == int(host.meta[:, 1].sum())


def test_checkerboard_worst_case_property(mesher):
    """Config 4 at reduced count: every chunk of a world-parity checkerboard block emits exactly
    98 304 quads in both modes (interior chunks see the pattern continue across all seams)."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table, PATTERN_CHECKER
    coords = grid_coords((0, 0, 0), (3, 3, 3))
    nbr = neighbour_table(coords)
    dvox = mesher.pattern_fill(dev(coords), PATTERN_CHECKER)
    for mode, _ in MODES:
        out = mesher.mesh(dvox, dev(nbr), mode, capacity=27 * 98304)
        torch.cuda.synchronize()
        host = out.to_host()
        assert (host.meta[:, 1] == 98304).all() and host.total_quads == 27 * 98304
        keys = S.keys_from_verts(host.chunk(13)[0])                  # the interior chunk
        assert len(np.unique(keys)) == 98304
        assert ((keys >> np.uint64(21)) & np.uint64(1023) == 0).all()   # all 1x1
        vox = dvox.cpu().numpy().view(np.uint16)
        nbs = [vox[nbr[13, f]] for f in range(6)]
        assert np.array_equal(np.sort(keys), O.mesh_chunk_keys(vox[13], nbs, mode))


def test_capacity_overflow_contract(G, mesher):
    vox, nbr = G["terrain42/vox"], G["terrain42/nbr"]
    need = int(G["terrain42/greedy/offs"][-1])
    import torch
    from voxelphile_b200 import MeshError
    out = mesher.mesh(dev(vox.view(np.int16)), dev(nbr), 1, capacity=need // 3)
    torch.cuda.synchronize()
    assert int(out.total.item()) == need
    meta = out.meta.cpu().numpy().view(np.uint32)
    lost = meta[:, 0] == 0xFFFFFFFF
    assert lost.any() and not lost.all()
    counts = np.diff(G["terrain42/greedy/offs"])
    assert np.array_equal(meta[:, 1], counts)
    with pytest.raises(MeshError) as e:
        out.to_host()
    assert e.value.kind == "Capacity"
    # exact capacity fits
    out = mesher.mesh(dev(vox.view(np.int16)), dev(nbr), 1, capacity=need)
    torch.cuda.synchronize()
    check_against_oracle(out.to_host(), vox, nbr, 1)


def test_host_level_api(G, mesher):
    """vxp_mesh_chunks_host / vxp_terrain_mesh_host / vxp_terrain_fill_host: host buffers in and out,
    including the internal arena growth (random_p50 needs more than the first-guess capacity)."""
    for name in ("terrain42", "random_p50", "empty"):
        vox, nbr = G[f"{name}/vox"], G[f"{name}/nbr"]
        for mode, _ in MODES:
            host = mesher.mesh_host(vox, nbr, mode)
            check_against_oracle(host, vox, nbr, mode)
    coords = G["terrain_coords"]
    assert np.array_equal(mesher.terrain_fill_host(coords, 42), G["terrain_vox"])
    host = mesher.terrain_mesh_host(np.array([[0, 8, 0]], np.int32), 42, 1)
    vox = O.terrain_fill(42, 0, 8, 0)
    nbs = [O.terrain_fill(42, d[0], 8 + d[1], d[2]) for d in S.DIRS]
    assert np.array_equal(np.sort(S.keys_from_verts(host.chunk(0)[0])), O.mesh_chunk_keys(vox, nbs, 1))
    empty = mesher.mesh_host(np.zeros((0, 32768), np.uint16), None, 1)
    assert empty.total_quads == 0 and empty.meta.shape == (0, 2)


def test_bad_arguments(mesher):
    from voxelphile_b200 import MeshError
    vox = np.zeros((2, 32768), np.uint16)
    with pytest.raises(MeshError) as e:
        mesher.mesh_host(vox, np.array([[5, -1, -1, -1, -1, -1], [-1] * 6], np.int32), 1)
    assert e.value.kind == "BadArg"
    with pytest.raises(MeshError):
        mesher.mesh_host(vox, None, 7)


def test_sharded_union_equals_single(mesher):
    """Multi-GPU semantics on one device: meshing two shards (each with the one-chunk seam layer
    duplicated read-only) yields the same per-chunk canonical meshes as meshing the whole batch."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table
    from voxelphile_b200.shard import shard_with_halo
    coords = grid_coords((0, 6, 0), (4, 10, 4))
    nbr = neighbour_table(coords)
    dvox = mesher.terrain_fill(dev(coords), 42)
    torch.cuda.synchronize()
    vox = dvox.cpu().numpy().view(np.uint16)
    whole = gpu_mesh(mesher, vox, nbr, 1, capacity=len(coords) * 4096)
    for world in (2, 3):
        seen = 0
        for rank in range(world):
            sh = shard_with_halo(coords, rank, world)
            part = gpu_mesh(mesher, vox[sh.batch_index], sh.nbr6, 1, capacity=len(sh.batch_index) * 4096,
                            n_mesh=sh.n_owned)
            for local, g in enumerate(sh.batch_index[:sh.n_owned]):
                a = np.sort(S.keys_from_verts(part.chunk(local)[0]))
                b = np.sort(S.keys_from_verts(whole.chunk(int(g))[0]))
                assert np.array_equal(a, b), (world, rank, int(g))
                seen += 1
        assert seen == len(coords)


def _chunk_hashes(host):
    """Order-independent hash (SPEC §4) of every chunk of a HostMesh, vectorised."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    keys = S.keys_from_verts(host.verts.reshape(-1, 4)) if host.total_quads else np.zeros(0, np.uint64)
    with np.errstate(over="ignore"):
        z = keys + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    out = np.zeros(host.meta.shape[0], dtype=np.uint64)
    owner = np.zeros(host.total_quads, dtype=np.int64)
    first, cnt = host.meta[:, 0].astype(np.int64), host.meta[:, 1].astype(np.int64)
    for i in np.nonzero(cnt)[0]:
        owner[first[i]:first[i] + cnt[i]] = i
    with np.errstate(over="ignore"):
        np.add.at(out, owner, z)
    return out & M


@pytest.mark.parametrize("mode,mname", MODES)
@pytest.mark.parametrize("seed", [42, 43, 44, 45])
def test_full_size_batch_hash_parity(mesher, mode, mname, seed):
    """BASELINE configs 2 / 3 at full size (4096 terrain chunks, coords [0,16)^3), for every seed bench.py times
    (42..45): per-chunk quad counts and order-independent key hashes equal the oracle's for every chunk, through
    the device-level call and (seed 42) through the sliced host-level call."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table
    coords = grid_coords((0, 0, 0), (16, 16, 16))
    nbr = neighbour_table(coords)
    dvox = mesher.terrain_fill(dev(coords), seed)
    torch.cuda.synchronize()
    vox = dvox.cpu().numpy().view(np.uint16)
    want_hash, want_cnt = O.mesh_batch_hash(vox, nbr, mode)
    out = mesher.mesh(dvox, dev(nbr), mode, capacity=int(want_cnt.sum()))
    torch.cuda.synchronize()
    for host in (out.to_host(),) + ((mesher.mesh_host(vox, nbr, mode),) if seed == 42 else ()):
        assert host.total_quads == int(want_cnt.sum())
        assert np.array_equal(host.meta[:, 1], want_cnt)
        assert np.array_equal(_chunk_hashes(host), want_hash)
        order = np.argsort(host.meta[:, 0][want_cnt > 0])
        f, c = host.meta[:, 0][want_cnt > 0][order].astype(np.int64), want_cnt[want_cnt > 0][order].astype(np.int64)
        assert f[0] == 0 and np.array_equal(f[1:], np.cumsum(c)[:-1]), "ranges must tile [0, Q)"


@pytest.mark.parametrize("mode,mname", MODES)
def test_permuted_chunk_order(mesher, mode, mname):
    """The neighbour table is arbitrary: a terrain block stored in shuffled order (so that x-neighbours are
    nowhere near each other in the batch) meshes to the same per-chunk result."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table
    rng = np.random.default_rng(5)
    coords = grid_coords((0, 6, 0), (6, 10, 5))
    coords = coords[rng.permutation(coords.shape[0])]
    nbr = neighbour_table(coords)
    dvox = mesher.terrain_fill(dev(coords), 42)
    torch.cuda.synchronize()
    vox = dvox.cpu().numpy().view(np.uint16)
    want_hash, want_cnt = O.mesh_batch_hash(vox, nbr, mode)
    host = gpu_mesh(mesher, vox, nbr, mode, capacity=int(want_cnt.sum()))
    assert np.array_equal(host.meta[:, 1], want_cnt)
    assert np.array_equal(_chunk_hashes(host), want_hash)


def test_host_pipeline_with_halo_chunks(mesher):
    """vxp_mesh_chunks_host in slices (n >= 256) with read-only halo chunks behind the meshed ones and a
    neighbour table that reaches across slices."""
    from voxelphile_b200 import grid_coords, neighbour_table
    from voxelphile_b200.shard import shard_with_halo
    coords = grid_coords((0, 6, 0), (12, 10, 12))             # 576 chunks
    vox = O.terrain_fill_batch(42, coords)
    sh = shard_with_halo(coords, 0, 2)                        # first half owned, seam layer appended
    want_hash, want_cnt = O.mesh_batch_hash(vox, neighbour_table(coords), 1)
    host = mesher.mesh_host(vox[sh.batch_index], sh.nbr6, 1, n_mesh=sh.n_owned)
    assert sh.n_owned >= 256 and len(sh.batch_index) > sh.n_owned
    assert np.array_equal(host.meta[:, 1], want_cnt[:sh.n_owned])
    assert np.array_equal(_chunk_hashes(host), want_hash[:sh.n_owned])


@pytest.mark.parametrize("mode,mname", MODES)
def test_overlapped_launches(mode, mname):
    """vxp_ctx_set_overlap: a stream of back-to-back mesh calls over four different full-size batches, two
    output meshes used alternately, gives every call the result of the same call made alone (per-chunk counts,
    ranges and key hashes vs the oracle for one batch; bit-identical per-chunk hashes vs serial for all)."""
    import torch
    from voxelphile_b200 import Mesher, grid_coords, neighbour_table
    m = Mesher(0)
    m.use_current_torch_stream()
    try:
        coords = grid_coords((0, 0, 0), (16, 16, 16))
        nbr = neighbour_table(coords)
        dn = dev(nbr)
        batches = [m.terrain_fill(dev(coords), 42 + k) for k in range(4)]
        serial = []
        for b in batches:
            o = m.mesh(b, dn, mode, capacity=4096 * 700)
            torch.cuda.synchronize()
            h = o.to_host()
            serial.append((h.total_quads, h.meta[:, 1].copy(), _chunk_hashes(h)))
        vox0 = batches[0].cpu().numpy().view(np.uint16)
        want_hash, want_cnt = O.mesh_batch_hash(vox0, nbr, mode)
        assert np.array_equal(serial[0][1], want_cnt) and np.array_equal(serial[0][2], want_hash)
        outs = [m.alloc_mesh(4096, 4096 * 700), m.alloc_mesh(4096, 4096 * 700)]
        m.set_overlap(True)
        torch.cuda.synchronize()
        for r in range(8):
            # 2r+1 calls in a row, then look at the last two (still intact: call k+2 is what re-uses call k's arrays)
            k_last = 2 * r + 1
            for k in range(k_last + 1):
                m.mesh(batches[(k + r) % 4], dn, mode, out=outs[k % 2])
            torch.cuda.synchronize()
            for k in (k_last - 1, k_last):
                h = outs[k % 2].to_host()
                tq, cnt, hs = serial[(k + r) % 4]
                assert h.total_quads == tq, (r, k)
                assert np.array_equal(h.meta[:, 1], cnt), (r, k)
                assert np.array_equal(_chunk_hashes(h), hs), (r, k)
                order = np.argsort(h.meta[:, 0][cnt > 0])
                f, c = h.meta[:, 0][cnt > 0][order].astype(np.int64), cnt[cnt > 0][order].astype(np.int64)
                assert f[0] == 0 and np.array_equal(f[1:], np.cumsum(c)[:-1]), "ranges must tile [0, Q)"
        # a long stream (what bench.py does): the last two results are still exact
        for k in range(400):
            m.mesh(batches[k % 4], dn, mode, out=outs[k % 2])
        torch.cuda.synchronize()
        for k in (398, 399):
            h = outs[k % 2].to_host()
            tq, cnt, hs = serial[k % 4]
            assert h.total_quads == tq and np.array_equal(h.meta[:, 1], cnt) and np.array_equal(_chunk_hashes(h), hs), k
        # modes and other calls may be mixed freely: a fill between two calls serialises them
        other = 1 - mode
        o1 = m.mesh(batches[1], dn, other, out=outs[0])
        scratch = m.terrain_fill(dev(coords), 42)
        o2 = m.mesh(scratch, dn, mode, out=outs[1])
        torch.cuda.synchronize()
        h2 = o2.to_host()
        assert h2.total_quads == serial[0][0] and np.array_equal(_chunk_hashes(h2), serial[0][2])
        m.set_overlap(False)
    finally:
        m.close()


# --------------------------------------------------------------------------- SPEC §9: incremental re-mesh
def _edit_world():
    from voxelphile_b200 import grid_coords, neighbour_table
    coords = grid_coords((0, 6, 0), (4, 10, 4))               # 64 chunks around the terrain surface
    return coords, neighbour_table(coords), O.terrain_fill_batch(42, coords)


def _random_edits(rng, n, m):
    """m edits of distinct voxels; a third of them pinned to chunk faces / edges / corners."""
    from voxelphile_b200 import EDIT_DTYPE
    seen, rows = set(), []
    while len(rows) < m:
        c = int(rng.integers(0, n))
        xyz = [int(v) for v in rng.integers(0, 32, 3)]
        if rng.random() < 0.35:
            for a in rng.choice(3, size=int(rng.integers(1, 4)), replace=False):
                xyz[a] = int(rng.choice([0, 31]))
        if (c, *xyz) in seen:
            continue
        seen.add((c, *xyz))
        rows.append((c, *xyz, int(rng.choice([0, 0, 1, 2, 3, 9, 0x7FFF]))))
    return np.array(rows, dtype=EDIT_DTYPE), rows


def _check_remeshed(host, dirty, vox, nbr, mode):
    for k, c in enumerate(dirty):
        nbs = [None if nbr[c, f] < 0 else vox[nbr[c, f]] for f in range(6)]
        want = O.mesh_chunk_keys(vox[c], nbs, mode)
        assert int(host.meta[k, 1]) == len(want), (k, c)
        got = np.sort(S.keys_from_verts(host.chunk(k)[0])) if len(want) else np.zeros(0, np.uint64)
        assert np.array_equal(got, want), (k, c)
    assert host.total_quads == int(host.meta[:len(dirty), 1].sum())


@pytest.mark.parametrize("mode,mname", MODES)
@pytest.mark.parametrize("m", [1, 40, 500])
def test_remesh_edits_device(mesher, mode, mname, m):
    import torch
    coords, nbr, vox = _edit_world()
    rng = np.random.default_rng(100 + m + mode)
    edits, rows = _random_edits(rng, len(coords), m)
    dvox, dn = dev(vox.view(np.int16)), dev(nbr)
    dirty_t, count_t, out = mesher.remesh_edits(dvox, dn, dev(edits.view(np.int32).reshape(-1, 3)), mode)
    torch.cuda.synchronize()
    want_dirty = S.apply_edits(vox, nbr, rows)
    cnt = int(count_t.item())
    dirty = dirty_t[:cnt].cpu().numpy().astype(np.int64)
    assert sorted(dirty.tolist()) == want_dirty and len(set(dirty.tolist())) == cnt
    assert np.array_equal(dvox.cpu().numpy().view(np.uint16), vox), "voxels after the edits"
    _check_remeshed(out.to_host(), dirty, vox, nbr, mode)


def test_mesh_list(mesher):
    """An explicit list: arbitrary order, duplicates allowed, meta entry k belongs to list[k]."""
    import torch
    coords, nbr, vox = _edit_world()
    lst = np.array([5, 63, 5, 0, 21, 22, 37], dtype=np.int32)
    for mode, _ in MODES:
        out = mesher.mesh_list(dev(vox.view(np.int16)), dev(nbr), dev(lst), mode)
        torch.cuda.synchronize()
        _check_remeshed(out.to_host(), lst, vox, nbr, mode)
    empty = mesher.mesh_list(dev(vox.view(np.int16)), dev(nbr), dev(np.zeros(0, np.int32)), 1)
    torch.cuda.synchronize()
    assert int(empty.total.item()) == 0


def test_remesh_edits_host(mesher):
    """Host-level flow: upload and mesh a world once, then send edits (12 bytes each) and get back only the
    re-meshed chunks; several rounds of edits accumulate in the resident world."""
    from voxelphile_b200 import EDIT_DTYPE, MeshError
    coords, nbr, vox = _edit_world()
    rng = np.random.default_rng(7)
    full = mesher.mesh_host(vox, nbr, 1)
    for rnd, m in enumerate((3, 60, 1)):
        edits, rows = _random_edits(rng, len(coords), m)
        dirty, host = mesher.remesh_edits_host(edits, rnd % 2)
        want_dirty = S.apply_edits(vox, nbr, rows)
        assert sorted(dirty.tolist()) == want_dirty
        _check_remeshed(host, dirty.astype(np.int64), vox, nbr, rnd % 2)
    # the resident world now equals the edited host copy: a full re-mesh of it agrees with the oracle
    again = mesher.mesh_host(vox, nbr, 1)
    want_hash, want_cnt = O.mesh_batch_hash(vox, nbr, 1)
    assert np.array_equal(again.meta[:, 1], want_cnt) and np.array_equal(_chunk_hashes(again), want_hash)
    assert full.total_quads != again.total_quads or not np.array_equal(full.meta, again.meta)
    dirty, host = mesher.remesh_edits_host(np.zeros(0, EDIT_DTYPE), 1)
    assert len(dirty) == 0 and host.total_quads == 0
    with pytest.raises(MeshError) as e:
        mesher.remesh_edits_host(np.array([(64, 0, 0, 0, 1)], dtype=EDIT_DTYPE), 1)
    assert e.value.kind == "BadArg"
    with pytest.raises(MeshError):
        mesher.remesh_edits_host(np.array([(0, 32, 0, 0, 1)], dtype=EDIT_DTYPE), 1)


def test_remesh_needs_a_resident_world():
    from voxelphile_b200 import EDIT_DTYPE, Mesher, MeshError
    m = Mesher(0)
    try:
        with pytest.raises(MeshError) as e:
            m.remesh_edits_host(np.array([(0, 1, 1, 1, 1)], dtype=EDIT_DTYPE), 1)
        assert e.value.kind == "BadArg"
    finally:
        m.close()


# --------------------------------------------------------------------------- SPEC §10: chunk wire format
def _palette_chunks(rng):
    """Chunks with 1, 2, 3, 4, 5, 16 and 17+ distinct ids (every index width), terrain chunks among them."""
    chunks = []
    for k in (1, 2, 3, 4, 5, 9, 16, 17, 300):
        ids = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 65535), size=k - 1, replace=False))]).astype(np.uint16)
        v = ids[rng.integers(0, k, 32768)]
        v[:k] = ids
        chunks.append(v)
    chunks.append(np.full(32768, 0xFFFF, np.uint16))
    from voxelphile_b200 import grid_coords
    chunks.extend(O.terrain_fill_batch(42, grid_coords((0, 6, 0), (2, 10, 2))))
    return np.stack(chunks)


def test_pack_matches_restatement_and_round_trips(mesher):
    import torch
    rng = np.random.default_rng(77)
    vox = _palette_chunks(rng)
    n = vox.shape[0]
    dv = dev(vox.view(np.int16))
    packed, offsets, total = mesher.pack(dv)
    back = mesher.unpack(packed, offsets)
    mesher.synchronize()
    torch.cuda.synchronize()
    hp, ho, ht = packed.cpu().numpy(), offsets.cpu().numpy().astype(np.int64), int(total.item())
    want = [S.pack_chunk(v) for v in vox]
    assert ht == sum(len(r) for r in want)
    spans = sorted((int(o), len(r)) for o, r in zip(ho, want))
    assert spans[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:])), "records must tile the buffer"
    for i in range(n):
        assert hp[ho[i]:ho[i] + len(want[i])].tobytes() == want[i], f"record {i}"
    assert np.array_equal(back.cpu().numpy().view(np.uint16), vox)
    # capacity contract: what does not fit is marked, the total still says how much is needed
    packed2, offsets2, total2 = mesher.pack(dv, capacity_bytes=4096)
    torch.cuda.synchronize()
    assert int(total2.item()) == ht
    o2 = offsets2.cpu().numpy().view(np.uint64)
    assert (o2 == np.uint64(0xFFFFFFFFFFFFFFFF)).any() and all(int(o) + len(want[i]) <= 4096 for i, o in enumerate(o2) if o != np.uint64(0xFFFFFFFFFFFFFFFF))


def test_unpack_accepts_any_palette_order_and_rejects_malformed(mesher):
    import torch
    from voxelphile_b200 import MeshError
    vox = np.zeros(32768, np.uint16); vox[::3] = 5; vox[1::7] = 9
    rec = bytearray(S.pack_chunk(vox))                       # palette [0, 5, 9], 2 bits
    # swap palette entries 1 and 2 and the indices with them: same chunk, non-canonical record
    pal = np.frombuffer(bytes(rec[16:48]), dtype="<u2").copy(); pal[[1, 2]] = pal[[2, 1]]
    data = np.frombuffer(bytes(rec[48:]), dtype=np.uint8)
    idx = np.stack([(data >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    idx = np.where(idx == 1, 2, np.where(idx == 2, 1, idx)).reshape(-1, 4)
    data2 = (idx[:, 0] | idx[:, 1] << 2 | idx[:, 2] << 4 | idx[:, 3] << 6).astype(np.uint8)
    rec2 = bytes(rec[:16]) + pal.tobytes() + data2.tobytes()
    buf = np.frombuffer(rec2, dtype=np.uint8).copy()
    got = mesher.unpack(dev(buf), dev(np.zeros(1, np.int64)))
    mesher.synchronize()
    assert np.array_equal(got.cpu().numpy().view(np.uint16).reshape(-1), vox)
    bad = buf.copy(); bad[4] = 3                               # bits = 3 is not a width of the format
    mesher.unpack(dev(bad), dev(np.zeros(1, np.int64)))
    with pytest.raises(MeshError) as e:
        mesher.synchronize()
    assert e.value.kind == "BadArg"
    mesher.synchronize()                                       # the verdict is reported once
    # calls queue without a synchronisation between them; a malformed record in any of them fails the next synchronize
    d_bad, d_ok, zero = dev(bad), dev(buf), dev(np.zeros(1, np.int64))
    mesher.unpack(d_ok, zero); mesher.unpack(d_bad, zero); last = mesher.unpack(d_ok, zero)
    with pytest.raises(MeshError):
        mesher.synchronize()
    assert np.array_equal(last.cpu().numpy().view(np.uint16).reshape(-1), vox)
    mesher.unpack(d_ok, zero); mesher.synchronize()
    with pytest.raises(MeshError):
        mesher.mesh_packed_host(bad, np.zeros(1, np.uint64), None, 1)
    # offsets that do not lie inside the buffer - the UINT64_MAX that vxp_pack_chunks gives a record that did not fit,
    # one past the end, a misaligned one, a record whose data runs past the end - are malformed records, not wild reads
    two = dev(np.concatenate([buf, buf]))
    for off in (0xFFFFFFFFFFFFFFFF, 2 * len(buf), 8, len(buf) + 16):
        offs = np.array([0, off], dtype=np.uint64)
        got = mesher.unpack(two[: len(buf) + 32] if off == len(buf) + 16 else two, dev(offs.view(np.int64)))
        with pytest.raises(MeshError):
            mesher.synchronize()
        g = got.cpu().numpy().view(np.uint16)
        assert np.array_equal(g[0], vox) and not g[1].any(), hex(off)
    # ... and packing straight into unpacking after a capacity overflow is such a case
    vv = np.stack([vox, vox, vox])
    pk, po, pt = mesher.pack(dev(vv.view(np.int16)), capacity_bytes=2 * len(buf))
    mesher.synchronize()
    assert int(pt.item()) == 3 * len(buf)
    back = mesher.unpack(pk, po)
    with pytest.raises(MeshError):
        mesher.synchronize()
    lost = po.cpu().numpy().view(np.uint64) == np.uint64(0xFFFFFFFFFFFFFFFF)
    assert lost.sum() == 1
    b = back.cpu().numpy().view(np.uint16)
    for i in range(3):
        assert (not b[i].any()) if lost[i] else np.array_equal(b[i], vox)


@pytest.mark.parametrize("mode,mname", MODES)
def test_mesh_packed_host_equals_mesh_host(mesher, mode, mname):
    """Packed upload path: same meshes as the raw upload path, and the world stays resident for edits."""
    from voxelphile_b200 import grid_coords, neighbour_table
    coords = grid_coords((0, 5, 0), (5, 11, 5))               # 150 chunks: air, surface and buried ones
    nbr = neighbour_table(coords)
    vox = O.terrain_fill_batch(42, coords)
    recs = [S.pack_chunk(v) for v in vox]
    order = np.random.default_rng(1).permutation(len(recs))   # records in arbitrary order
    offsets = np.zeros(len(recs), np.uint64)
    pos = 0
    for i in order:
        offsets[i] = pos
        pos += len(recs[i])
    buf = np.zeros(pos, np.uint8)
    for i in range(len(recs)):
        buf[int(offsets[i]):int(offsets[i]) + len(recs[i])] = np.frombuffer(recs[i], np.uint8)
    assert pos < vox.nbytes / 4
    a = mesher.mesh_packed_host(buf, offsets, nbr, mode)
    want_hash, want_cnt = O.mesh_batch_hash(vox, nbr, mode)
    assert np.array_equal(a.meta[:, 1], want_cnt) and np.array_equal(_chunk_hashes(a), want_hash)
    edits, rows = _random_edits(np.random.default_rng(3), len(coords), 10)
    dirty, host = mesher.remesh_edits_host(edits, mode)
    assert sorted(dirty.tolist()) == S.apply_edits(vox, nbr, rows)
    _check_remeshed(host, dirty.astype(np.int64), vox, nbr, mode)


def test_fused_world_full_size(mesher):
    """BASELINE config 5 at full size: the 65^3 = 274 625 chunk coordinates centred on the origin, seed 42, fused
    terrain + greedy: per-chunk quad counts and key hashes of EVERY chunk against the oracle, which generates and
    meshes the world one z-layer at a time (vxo_terrain_world_hash)."""
    import torch
    from voxelphile_b200 import grid_coords
    lo, hi = (-32, -32, -32), (33, 33, 33)
    coords = grid_coords(lo, hi)
    assert coords.shape[0] == 274625
    want_hash, want_cnt = O.terrain_world_hash(42, lo, hi, 1)
    out = mesher.terrain_mesh(dev(coords), 42, 1, capacity=int(want_cnt.sum()))
    torch.cuda.synchronize()
    host = out.to_host()
    assert host.total_quads == int(want_cnt.sum())
    assert np.array_equal(host.meta[:, 1], want_cnt)
    assert np.array_equal(_chunk_hashes(host), want_hash)
    # a second call on the same context (next key epoch of the column table, counters left zeroed by the first)
    out2 = mesher.terrain_mesh(dev(coords), 42, 1, capacity=int(want_cnt.sum()))
    torch.cuda.synchronize()
    host2 = out2.to_host()
    assert np.array_equal(host2.meta[:, 1], want_cnt) and np.array_equal(_chunk_hashes(host2), want_hash)


def test_fused_world_slab_culled_and_shuffled(mesher):
    """Six z-layers of the config-5 world (25 350 chunks) through the fused culled kernel, and the same chunks in a
    shuffled order through the fused greedy kernel (columns are found through the hash table, not by position)."""
    import torch
    from voxelphile_b200 import grid_coords
    lo, hi = (-32, -32, -3), (33, 33, 3)
    coords = grid_coords(lo, hi)
    for mode in (0, 1):
        want_hash, want_cnt = O.terrain_world_hash(42, lo, hi, mode)
        perm = np.random.default_rng(5).permutation(coords.shape[0]) if mode == 1 else np.arange(coords.shape[0])
        out = mesher.terrain_mesh(dev(coords[perm]), 42, mode, capacity=int(want_cnt.sum()))
        torch.cuda.synchronize()
        host = out.to_host()
        assert np.array_equal(host.meta[:, 1], want_cnt[perm]), mode
        assert np.array_equal(_chunk_hashes(host), want_hash[perm]), mode


def _device_chunk_hashes(out, n):
    """_chunk_hashes on the device (torch as the checker's arithmetic) for meshes too large to bring to the host:
    keys from vertex 0 of every quad (SPEC §5 inverse), mix64, per-chunk sums through a cumulative sum over the
    chunk ranges.  Returns (hashes u64 numpy, counts, ranges_tile)."""
    import torch
    Q = int(out.total.item())
    meta = out.meta.cpu().numpy().view(np.uint32).reshape(n, 2)
    first, cnt = meta[:, 0].astype(np.int64), meta[:, 1].astype(np.int64)
    v0 = out.verts[: 4 * Q].view(-1, 4)[:, 0]
    w0, w1 = v0 & 0xFFFFFFFF, (v0 >> 32) & 0xFFFFFFFF
    x, y, z, f = w0 & 63, (w0 >> 6) & 63, (w0 >> 12) & 63, (w0 >> 18) & 7
    axis = f >> 1
    along = torch.where(axis == 0, x, torch.where(axis == 1, y, z)) - (f & 1)
    u = torch.where(axis == 0, y, x)
    v = torch.where(axis == 2, y, z)
    key = u | v << 6 | along << 12 | f << 18 | ((w1 >> 16) & 31) << 21 | ((w1 >> 21) & 31) << 26 | (w1 & 0xFFFF) << 31
    del w0, w1, x, y, z, f, axis, along, u, v

    def shr(t, k):                                                   # logical shift of the u64 bit pattern held in int64
        return (t >> k) & ((1 << (64 - k)) - 1)

    def c64(c):                                                      # u64 constant as the int64 with the same bits
        return c - (1 << 64) if c >= (1 << 63) else c

    zt = key + c64(0x9E3779B97F4A7C15)
    zt = (zt ^ shr(zt, 30)) * c64(0xBF58476D1CE4E5B9)
    zt = (zt ^ shr(zt, 27)) * c64(0x94D049BB133111EB)
    zt = zt ^ shr(zt, 31)
    cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=zt.device), torch.cumsum(zt, 0)])
    fi = torch.from_numpy(first).to(zt.device)
    ci = torch.from_numpy(cnt).to(zt.device)
    fi = torch.where(ci > 0, fi, torch.zeros_like(fi))
    h = (cs[fi + ci] - cs[fi]).cpu().numpy().view(np.uint64)
    nz = cnt > 0
    order = np.argsort(first[nz])
    f_s, c_s = first[nz][order], cnt[nz][order]
    tiles = len(f_s) == 0 or (f_s[0] == 0 and np.array_equal(f_s[1:], np.cumsum(c_s)[:-1]) and f_s[-1] + c_s[-1] == Q)
    return h, cnt, tiles


@pytest.mark.parametrize("mode,mname", MODES)
def test_checkerboard_full_size(mesher, mode, mname):
    """BASELINE config 4 at the size bench.py times: 4096 chunks ([0,16)^3) of the world-parity checkerboard,
    402 653 184 quads (22.5 GB of mesh).  Every chunk: 98 304 quads and the oracle's key hash; the ranges tile
    [0, Q); the indices of a sampled range are the chunk-local ordinals."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table, PATTERN_CHECKER
    coords = grid_coords((0, 0, 0), (16, 16, 16))
    nbr = neighbour_table(coords)
    dvox = mesher.pattern_fill(dev(coords), PATTERN_CHECKER)
    torch.cuda.synchronize()
    vox = dvox.cpu().numpy().view(np.uint16)
    want_hash, want_cnt = O.mesh_batch_hash(vox, nbr, mode)
    assert (want_cnt == 98304).all()
    out = mesher.mesh(dvox, dev(nbr), mode, capacity=4096 * 98304)
    torch.cuda.synchronize()
    assert int(out.total.item()) == 4096 * 98304
    h, cnt, tiles = _device_chunk_hashes(out, 4096)
    assert (cnt == 98304).all() and tiles
    assert np.array_equal(h, want_hash)
    first = int(out.meta[1234, 0].item())
    ix = out.indices[6 * first: 6 * (first + 98304)].cpu().numpy().view(np.uint32).reshape(-1, 6)
    assert np.array_equal(ix, 4 * np.arange(98304, dtype=np.uint32)[:, None] + np.array([0, 1, 2, 2, 3, 0], dtype=np.uint32))
    del out
    torch.cuda.empty_cache()


def test_greedy_staging_fallbacks(mesher):
    """Chunks that do not fit the greedy kernel's staging area in one go: sparse layers of per-voxel random ids
    (few non-empty plane rows, thousands of quads: staged and flushed in windows of ordinals) and dense noise
    (the sweep log itself does not fit: stepped sweep)."""
    rng = np.random.default_rng(21)
    v = np.zeros((3, 32, 32, 32), dtype=np.uint16)                   # [z][y][x]
    v[0, 2::6] = rng.integers(1, 5, (5, 32, 32))                     # five separate z-layers
    v[1, :, 3::7, :] = rng.integers(1, 7, (32, 5, 32))               # five separate y-layers
    v[2] = np.where(rng.random((32, 32, 32)) < 0.5, rng.integers(1, 5, (32, 32, 32)), 0)
    vox = v.reshape(3, -1)
    host = gpu_mesh(mesher, vox, None, 1)
    assert host.meta[0, 1] > 7000 and host.meta[1, 1] > 7000
    check_against_oracle(host, vox, None, 1)


def test_greedy_log_overflow_in_halves(mesher):
    """Chunks whose sweep log does not fit in one piece but does in two halves of the planes (between 1985 and ~2900
    non-empty plane rows): scattered single voxels at 1.2 % .. 2 % density, one and several ids, with and without a
    neighbour; denser ones take the stepped sweep (covered above)."""
    rng = np.random.default_rng(77)
    dens = (0.012, 0.014, 0.016, 0.018, 0.020, 0.024)
    vox = np.zeros((2 * len(dens) + 1, 32768), dtype=np.uint16)
    for k, p in enumerate(dens):
        hit = rng.random(32768) < p
        vox[2 * k][hit] = 3
        vox[2 * k + 1] = np.where(rng.random(32768) < p, rng.integers(1, 6, 32768), 0)
    vox[-1] = np.where(rng.random(32768) < 0.3, 2, 0)                 # something to look at across +x
    nbr = np.full((vox.shape[0], 6), -1, dtype=np.int32)
    nbr[:-1, 1] = vox.shape[0] - 1
    for mode in (1,):
        host = gpu_mesh(mesher, vox, nbr, mode)
        check_against_oracle(host, vox, nbr, mode)
    # the rows of such a chunk: every voxel is alone, so it has one row in six planes
    rows = [6 * int((vox[i] != 0).sum()) for i in range(len(vox) - 1)]
    assert min(rows) < 2944 and max(rows) > 1984


# --------------------------------------------------------------------------- SPEC §11: level of detail
def test_downsample_and_mesh_lod(mesher):
    """vxp_downsample_chunks against the numpy restatement (random chunks, missing children, terrain), and the
    downsampled terrain meshed by the ordinary kernels against the oracle mesh of the restated LOD chunks."""
    import torch
    from voxelphile_b200 import grid_coords, neighbour_table
    rng = np.random.default_rng(9)
    vox = np.where(rng.random((20, 32768)) < 0.4, rng.integers(1, 400, (20, 32768)), 0).astype(np.uint16)
    child = rng.integers(-1, 20, size=(6, 8)).astype(np.int32)
    child[0] = -1
    got = mesher.downsample(dev(vox.view(np.int16)), dev(child))
    torch.cuda.synchronize()
    got = got.cpu().numpy().view(np.uint16)
    for i in range(6):
        want = S.downsample([None if c < 0 else vox[c] for c in child[i]])
        assert np.array_equal(got[i], want), i
    # two levels of a terrain block: 8x8x8 chunks -> 4x4x4 -> 2x2x2, each level meshed greedy
    coords = grid_coords((0, 4, 0), (8, 12, 8))
    level = O.terrain_fill_batch(42, coords)
    dlevel = dev(level.view(np.int16))
    dim = 8
    while dim > 2:
        half = dim // 2
        idx = lambda x, y, z, d=dim: x + d * (y + d * z)             # grid_coords order: x fastest
        assert tuple(coords[idx(1, 0, 0)] - coords[0]) == (1, 0, 0) or dim != 8
        kids = np.array([[idx(2 * X + (c & 1), 2 * Y + ((c >> 1) & 1), 2 * Z + (c >> 2)) for c in range(8)]
                         for Z in range(half) for Y in range(half) for X in range(half)], dtype=np.int32)
        want = np.stack([S.downsample([level[c] for c in row]) for row in kids])
        dlevel = mesher.downsample(dlevel, dev(kids))
        torch.cuda.synchronize()
        assert np.array_equal(dlevel.cpu().numpy().view(np.uint16), want), dim
        lod_coords = grid_coords((0, 0, 0), (half, half, half))
        nbr = neighbour_table(lod_coords)
        out = mesher.mesh(dlevel, dev(nbr), 1, capacity=half ** 3 * 8192)
        torch.cuda.synchronize()
        want_hash, want_cnt = O.mesh_batch_hash(want, nbr, 1)
        host = out.to_host()
        assert np.array_equal(host.meta[:, 1], want_cnt) and np.array_equal(_chunk_hashes(host), want_hash), dim
        level, dim = want, half


def test_terrain_fill_large_batches_bit_exact(mesher):
    """Batches of >= 16 384 chunks take the shared-column path of vxp_terrain_fill_batch: one batch of 17 216 chunks
    made of a tall block (16 chunks per column), a flat sheet (every column its own) and a shuffled list with
    negative coordinates (some columns shared with the block), against the oracle; then smaller batches (the
    one-launch shared path) on the same context, and the large one again (next epoch of the column table)."""
    import torch
    from voxelphile_b200 import grid_coords
    rng = np.random.default_rng(13)
    tall = grid_coords((0, 0, 0), (24, 16, 24))                      # 9216 chunks, 576 columns
    sheet = grid_coords((-70, 8, -40), (-10, 9, 20))                 # 3600 chunks, 3600 columns
    mixed = grid_coords((-10, 2, -10), (10, 13, 10))                 # 4400 chunks
    mixed = mixed[rng.permutation(mixed.shape[0])]
    coords = np.concatenate([tall, sheet, mixed])
    assert coords.shape[0] >= 16384
    seed = 2**63 + 5
    want = O.terrain_fill_batch(seed, coords)
    for rep in range(2):
        got = mesher.terrain_fill(dev(coords), seed)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy().view(np.uint16), want), rep
        del got
        small = mesher.terrain_fill(dev(coords[:1000]), seed)
        torch.cuda.synchronize()
        assert np.array_equal(small.cpu().numpy().view(np.uint16), want[:1000])
        # 256 .. 16 383 chunks: one launch, the first chunk of a column computes its heights for the others
        # (stacked, flat and shuffled columns in one batch; twice in a row: next epoch of the table)
        pick = np.concatenate([np.arange(9216, 9216 + 1500), np.arange(12816, 12816 + 2500), np.arange(0, 2000)])
        for again in range(2):
            mid = mesher.terrain_fill(dev(coords[pick]), seed)
            torch.cuda.synchronize()
            assert np.array_equal(mid.cpu().numpy().view(np.uint16), want[pick]), (rep, again)
            del mid
    # the fused path right after the column fill (they share the column scratch and its launch counters)
    fc = grid_coords((0, 6, 0), (3, 10, 2))
    wh, wc = O.terrain_world_hash(42, (0, 6, 0), (3, 10, 2), 1)
    out = mesher.terrain_mesh(dev(fc), 42, 1, capacity=int(wc.sum()))
    torch.cuda.synchronize()
    host = out.to_host()
    assert np.array_equal(host.meta[:, 1], wc) and np.array_equal(_chunk_hashes(host), wh)


@pytest.mark.parametrize("mode,mname", MODES)
def test_meshes_without_an_index_buffer(G, mesher, mode, mname):
    """vxp_mesh_dev.indices == NULL / vxp_ctx_set_host_indices(0): same vertices, nothing written past them."""
    import torch
    vox, nbr = G["terrain42/vox"], G["terrain42/nbr"]
    want_h, want_c = O.mesh_batch_hash(vox, nbr, mode)
    out = mesher.alloc_mesh(vox.shape[0], int(want_c.sum()), indices=False)
    mesher.mesh(dev(vox.view(np.int16)), dev(nbr), mode, out=out)
    torch.cuda.synchronize()
    host = out.to_host()
    assert host.indices.size == 0 and np.array_equal(host.meta[:, 1], want_c) and np.array_equal(_chunk_hashes(host), want_h)
    mesher.set_host_indices(False)
    try:
        for h in (mesher.mesh_host(vox, nbr, mode), mesher.mesh_packed_host(*_packed(vox), nbr, mode)):
            assert h.indices.size == 0 and np.array_equal(h.meta[:, 1], want_c) and np.array_equal(_chunk_hashes(h), want_h)
    finally:
        mesher.set_host_indices(True)
    h = mesher.mesh_host(vox, nbr, mode)
    assert h.indices.size == 6 * h.total_quads


def _packed(vox):
    recs = [S.pack_chunk(v) for v in vox]
    offsets = np.cumsum([0] + [len(r) for r in recs[:-1]]).astype(np.uint64)
    return np.frombuffer(b"".join(recs), dtype=np.uint8).copy(), offsets


def test_default_stream_configuration(G):
    """A Mesher used as constructed (no use_current_torch_stream call): device tensors made on torch's stream, the
    kernels on the stream the Mesher chose, results read back through torch - ordering is the Mesher's job."""
    import torch
    from voxelphile_b200 import Mesher
    m = Mesher(0)
    vox, nbr = G["terrain42/vox"], G["terrain42/nbr"]
    want_h, want_c = O.mesh_batch_hash(vox, nbr, 1)
    side = torch.cuda.Stream()
    for rep in range(20):
        with torch.cuda.stream(side if rep % 2 else torch.cuda.default_stream()):
            dv = dev(vox.view(np.int16))
            dn = dev(nbr)
            junk = torch.full((1 << 22,), rep, dtype=torch.int64, device="cuda")   # keeps the allocating stream busy
            out = m.mesh(dv, dn, 1, capacity=int(want_c.sum()))
            host = out.to_host()
            del junk
        assert np.array_equal(host.meta[:, 1], want_c), rep
        assert np.array_equal(_chunk_hashes(host), want_h), rep
    m.close()


def test_new_entry_points_reject_bad_arguments_and_accept_empty_batches(mesher):
    """Through the raw C ABI: null pointers and empty batches for the re-mesh, wire-format and LOD calls."""
    import ctypes as C
    import torch
    from voxelphile_b200 import _lib as L
    lib, h = mesher._lib, mesher._h
    BAD, OK = L.VXP_ERR_BAD_ARG, L.VXP_OK
    vox = torch.zeros((2, 32768), dtype=torch.int16, device="cuda")
    out = mesher.alloc_mesh(2, 1024)
    s = mesher._mesh_struct(out)
    lst = torch.zeros(2, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    # vxp_mesh_list
    assert lib.vxp_mesh_list(h, vox.data_ptr(), None, 2, None, 2, 1, C.byref(s)) == BAD
    assert lib.vxp_mesh_list(h, None, None, 2, lst.data_ptr(), 2, 1, C.byref(s)) == BAD
    assert lib.vxp_mesh_list(h, vox.data_ptr(), None, 2, lst.data_ptr(), 2, 9, C.byref(s)) == BAD
    assert lib.vxp_mesh_list(h, vox.data_ptr(), None, 2, lst.data_ptr(), 0, 1, C.byref(s)) == OK
    # vxp_remesh_edits
    ed = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    assert lib.vxp_remesh_edits(h, vox.data_ptr(), None, 2, ed.data_ptr(), 1, 1, lst.data_ptr(), None, C.byref(s)) == BAD
    assert lib.vxp_remesh_edits(h, vox.data_ptr(), None, 2, None, 1, 1, lst.data_ptr(), cnt.data_ptr(), C.byref(s)) == BAD
    assert lib.vxp_remesh_edits(h, vox.data_ptr(), None, 2, ed.data_ptr(), 0, 1, lst.data_ptr(), cnt.data_ptr(), C.byref(s)) == OK
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0 and int(out.total.item()) == 0
    # out-of-range edits are ignored, in-range ones applied
    e2 = np.array([(5, 0, 0, 0, 9), (1, 40, 0, 0, 9), (1, 3, 4, 5, 9)], dtype=EDIT_DTYPE)
    d, c, o = mesher.remesh_edits(vox, None, dev(e2.view(np.int32).reshape(-1, 3)), 1)
    torch.cuda.synchronize()
    assert int(c.item()) == 1 and int(d[0].item()) == 1 and int(vox[1, 3 + 32 * 4 + 1024 * 5].item()) == 9
    assert int(o.total.item()) == 6
    # vxp_pack_chunks / vxp_unpack_chunks / vxp_downsample_chunks
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    off = torch.zeros(2, dtype=torch.int64, device="cuda")
    buf = torch.zeros(1 << 18, dtype=torch.uint8, device="cuda")
    assert lib.vxp_pack_chunks(h, vox.data_ptr(), 2, buf.data_ptr(), 1 << 18, off.data_ptr(), None) == BAD
    assert lib.vxp_pack_chunks(h, None, 2, buf.data_ptr(), 1 << 18, off.data_ptr(), tot.data_ptr()) == BAD
    assert lib.vxp_pack_chunks(h, vox.data_ptr(), 2, buf.data_ptr() + 8, 1 << 17, off.data_ptr(), tot.data_ptr()) == BAD
    assert lib.vxp_pack_chunks(h, vox.data_ptr(), 0, None, 0, None, tot.data_ptr()) == OK
    assert lib.vxp_unpack_chunks(h, None, 0, off.data_ptr(), 2, vox.data_ptr()) == BAD
    assert lib.vxp_unpack_chunks(h, buf.data_ptr(), buf.numel(), off.data_ptr(), 0, None) == OK
    kids = torch.full((1, 8), -1, dtype=torch.int32, device="cuda")
    assert lib.vxp_downsample_chunks(h, vox.data_ptr(), None, 1, vox.data_ptr()) == BAD
    assert lib.vxp_downsample_chunks(h, vox.data_ptr(), kids.data_ptr(), 0, None) == OK
    lod = mesher.downsample(vox, kids)
    torch.cuda.synchronize()
    assert int(lod.abs().sum().item()) == 0                       # eight missing children: air
    mesher.synchronize()
    # overlap switch is a plain flag
    assert lib.vxp_ctx_set_overlap(None, 1) == BAD and lib.vxp_ctx_set_overlap(h, 1) == OK and lib.vxp_ctx_set_overlap(h, 0) == OK
