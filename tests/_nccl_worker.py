"""Worker for tests/test_multi_gpu.py: one rank of a torchrun job, one B200 per rank, NCCL.
Every rank meshes its shard with the CUDA kernels (libvxp.so, through the C ABI); the per-chunk quad counts
and key hashes are gathered over NCCL (the path's one collective) and rank 0 compares them with the oracle's
mesh of the UNSHARDED world.  Two cases: a stored-voxel block sharded with read-only seam chunks (both modes),
and the fused terrain world strong-scaled by coordinate ranges."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import spec_py as S  # noqa: E402
from voxelphile_b200 import Mesher, grid_coords, neighbour_table  # noqa: E402
from voxelphile_b200.shard import gather_counts, partition, shard_with_halo  # noqa: E402


def chunk_hashes(host) -> np.ndarray:
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
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    m = Mesher(local)

    def gather(values_u64, n_total):
        t = torch.from_numpy(values_u64.astype(np.uint64).view(np.int64)).to(dev)
        return gather_counts(t, n_total, rank, world).cpu().numpy().view(np.uint64)

    # ---- stored voxels, sharded with read-only seam chunks
    coords = grid_coords((0, 4, 0), (8, 12, 8))                       # 512 chunks across the surface
    nbr = neighbour_table(coords)
    vox = O.terrain_fill_batch(42, coords)                            # every rank holds the world on the host
    sh = shard_with_halo(coords, rank, world, nbr)
    dvox = torch.from_numpy(vox[sh.batch_index].view(np.int16)).to(dev)
    dnbr = torch.from_numpy(sh.nbr6).to(dev)
    for mode in (0, 1):
        out = m.mesh(dvox, dnbr, mode, capacity=sh.n_owned * 8192, n_mesh=sh.n_owned)
        host = out.to_host()
        counts = gather(host.meta[:, 1].astype(np.uint64), len(coords))
        hashes = gather(chunk_hashes(host), len(coords))
        if rank == 0:
            want_h, want_c = O.mesh_batch_hash(vox, nbr, mode)
            assert np.array_equal(counts, want_c.astype(np.uint64)), f"mode {mode}: gathered counts differ"
            assert np.array_equal(hashes, want_h), f"mode {mode}: sharded meshes differ from the unsharded oracle mesh"

    # ---- fused terrain world, strong-scaled by coordinate ranges (no halo: terrain is a pure function)
    lo, hi = (-6, 4, -6), (7, 13, 7)
    wcoords = grid_coords(lo, hi)
    a, b = partition(len(wcoords), rank, world)
    out = m.terrain_mesh(torch.from_numpy(np.ascontiguousarray(wcoords[a:b])).to(dev), 42, 1, capacity=(b - a) * 4096)
    host = out.to_host()
    counts = gather(host.meta[:, 1].astype(np.uint64), len(wcoords))
    hashes = gather(chunk_hashes(host), len(wcoords))
    if rank == 0:
        want_h, want_c = O.terrain_world_hash(42, lo, hi, 1)
        assert np.array_equal(counts, want_c.astype(np.uint64)), "fused world: gathered counts differ"
        assert np.array_equal(hashes, want_h), "fused world: sharded meshes differ from the oracle"
        print("NCCL_OK", world, int(counts.sum()))
    dist.barrier()
    dist.destroy_process_group()
    m.close()


if __name__ == "__main__":
    main()
