"""Small launch sequence for ncu captures (GPU box): fills the four config-2/3 batches, then N launches of one kernel.
usage: python tools/ncu_target.py {greedy|culled|fused_greedy|fill|pack|unpack|lod} [N]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, CULLED, GREEDY
what = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
m = Mesher(0)
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
torch.cuda.synchronize()
if what in ("greedy", "culled"):
    mode = GREEDY if what == "greedy" else CULLED
    out = m.alloc_mesh(4096, 3_000_000)
    for k in range(N): m.mesh(vox[k % 4], dn, mode, out=out)
elif what == "fused_greedy":
    out = m.alloc_mesh(4096, 3_000_000)
    for k in range(N): m.terrain_mesh(dc, 42 + k % 4, GREEDY, out=out)
elif what == "fill":
    for k in range(N): m.terrain_fill(dc, 100 + k, out=vox[k % 4])
elif what == "pack":
    for k in range(N): m.pack(vox[k % 4])
elif what == "unpack":
    pk, po, pt = m.pack(vox[0]); out = torch.empty_like(vox[1])
    for k in range(N): m.unpack(pk, po, out=out)
elif what == "lod":
    k8 = np.array([[(2 * X + (c & 1)) + 16 * ((2 * Y + ((c >> 1) & 1)) + 16 * (2 * Z + (c >> 2))) for c in range(8)]
                   for Z in range(8) for Y in range(8) for X in range(8)], dtype=np.int32)
    dk8 = torch.from_numpy(k8).cuda(); out = torch.empty((512, 32768), dtype=torch.int16, device="cuda")
    for k in range(N): m.downsample(vox[k % 4], dk8, out=out)
torch.cuda.synchronize()
print("ok", what, N)
