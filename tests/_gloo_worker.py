"""Worker for tests/test_sharding.py: one rank of a world_size-N gloo job on CPU.
The oracle stands in for the GPU mesher so that the host-side sharding logic (partition,
seam halo chunks, local neighbour tables, the final gather of counts) runs exactly as in
bench.py --gpus N."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from voxelphile_b200 import grid_coords, neighbour_table  # noqa: E402
from voxelphile_b200.shard import gather_counts, shard_with_halo  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coords = grid_coords((0, 6, 0), (3, 10, 3))
    nbr = neighbour_table(coords)
    vox = O.terrain_fill_batch(42, coords, threads=1)
    sh = shard_with_halo(coords, rank, world, nbr)
    local_vox = vox[sh.batch_index]
    counts, hashes = [], []
    for i in range(sh.n_owned):
        nbs = [None if sh.nbr6[i, f] < 0 else local_vox[sh.nbr6[i, f]] for f in range(6)]
        k = O.mesh_chunk_keys(local_vox[i], nbs, 1)
        counts.append(len(k))
        hashes.append(int(O.load().vxo_hash_keys(k.ctypes.data, len(k)) & (2**63 - 1)))
    all_counts = gather_counts(torch.tensor(counts, dtype=torch.int64), len(coords), rank, world)
    all_hashes = gather_counts(torch.tensor(hashes, dtype=torch.int64), len(coords), rank, world)
    if rank == 0:
        want_h, want_c = O.mesh_batch_hash(vox, nbr, 1, threads=1)
        assert all_counts.tolist() == want_c.tolist(), "gathered counts differ from the unsharded batch"
        assert all_hashes.tolist() == [int(h) & (2**63 - 1) for h in want_h], "sharded meshes differ"
        print("GLOO_OK", int(all_counts.sum()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
