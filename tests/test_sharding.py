"""CPU tests of the multi-GPU host logic: partition, seam halo chunks and the
final gather of counts, including a real world_size-2 and -3 gloo job."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def test_partition_is_balanced_and_covers():
    from voxelphile_b200.shard import partition
    for n in (0, 1, 7, 4096, 274625):
        for world in (1, 2, 3, 4, 8):
            parts = [partition(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_halo_reproduces_unsharded_meshes():
    from voxelphile_b200 import grid_coords, neighbour_table
    from voxelphile_b200.shard import shard_with_halo
    coords = grid_coords((-1, 7, -1), (2, 9, 2))
    nbr = neighbour_table(coords)
    vox = O.terrain_fill_batch(7, coords)
    want_h, want_c = O.mesh_batch_hash(vox, nbr, 1)
    for world in (1, 2, 4, 5):
        seen = []
        for rank in range(world):
            sh = shard_with_halo(coords, rank, world, nbr)
            assert (sh.nbr6 < len(sh.batch_index)).all() and sh.nbr6.shape == (sh.n_owned, 6)
            assert len(set(sh.batch_index.tolist())) == len(sh.batch_index)
            h, c = O.mesh_batch_hash(vox[sh.batch_index], np.vstack([sh.nbr6, np.full((len(sh.batch_index) - sh.n_owned, 6), -1, np.int32)]), 1)
            for local in range(sh.n_owned):
                g = int(sh.batch_index[local])
                assert (h[local], c[local]) == (want_h[g], want_c[g])
                seen.append(g)
        assert sorted(seen) == list(range(len(coords)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_gather_of_counts(world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK" in outs[0]
