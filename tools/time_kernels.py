"""Kernel time of mesh variants (VXP_LIB) on the config-2 batches: culled and greedy, events around 100 launches."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, CULLED, GREEDY
m = Mesher(0); m.use_current_torch_stream()
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
res = []
for mode in (CULLED, GREEDY):
    out = m.alloc_mesh(4096, 4_000_000)
    for k in range(20): m.mesh(vox[k % 4], dn, mode, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(200): m.mesh(vox[k % 4], dn, mode, out=out)
    b.record(); torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 200 * 1e3)
print(f"{os.environ.get('VXP_LIB','libvxp.so').split('/')[-1]:22s} culled {res[0]:7.1f} us   greedy {res[1]:7.1f} us")
