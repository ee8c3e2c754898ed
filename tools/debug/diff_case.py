"""Debug helper (GPU box): reproduce test_differential_random for one parameter set and print the key differences."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O, spec_py as S
from voxelphile_b200 import Mesher
p, palette, mode = float(sys.argv[1]), sys.argv[2], int(sys.argv[3])
PALETTES = {"small": [1, 2, 3, 4], "large": [1000, 1001, 1002, 0xFFFF], "high_bits": [3, 11, 0x103, 0x8003]}
rng = np.random.default_rng(int(p * 1000) + mode)
n = 24 if palette == "small" else 8
ids = np.array(PALETTES[palette], dtype=np.uint16)
vox = np.where(rng.random((n, 32768)) < p, ids[rng.integers(0, 4, (n, 32768))], 0).astype(np.uint16)
nbr = rng.integers(-1, n, size=(n, 6)).astype(np.int32)
m = Mesher(0)
out = m.mesh(torch.from_numpy(vox.view(np.int16)).cuda(), torch.from_numpy(nbr).cuda(), mode, capacity=n * 98304)
host = out.to_host()
def dec(k):
    k = int(k); return dict(u=k & 63, v=(k >> 6) & 63, s=(k >> 12) & 63, f=(k >> 18) & 7, w=((k >> 21) & 31) + 1, h=((k >> 26) & 31) + 1, b=k >> 31)
for i in range(n):
    nbs = [None if nbr[i, f] < 0 else vox[nbr[i, f]] for f in range(6)]
    want = O.mesh_chunk_keys(vox[i], nbs, mode)
    got = np.sort(S.keys_from_verts(host.chunk(i)[0]))
    if not np.array_equal(got, want):
        extra = np.setdiff1d(got, want); missing = np.setdiff1d(want, got)
        print("chunk", i, "nbr", nbr[i].tolist(), "got", len(got), "want", len(want))
        for k in extra[:12]: print("   extra  ", dec(k))
        for k in missing[:12]: print("   missing", dec(k))
        u, c = np.unique(got, return_counts=True)
        print("   duplicates in got:", [dec(k) for k in u[c > 1][:6]])
print("done")
