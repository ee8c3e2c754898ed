"""GPU-side time of one culled batch (scan / mesh slices on two streams) when replayed as a CUDA graph,
i.e. without the host's launch overhead.  Timing only: a replayed graph re-uses the capture's epoch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, CULLED, GREEDY
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
m = Mesher(0); m.use_current_torch_stream()
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
for mode, name in ((CULLED, "culled"), (GREEDY, "greedy")):
    out = m.alloc_mesh(4096, 4_000_000)
    for k in range(8): m.mesh(vox[k % 4], dn, mode, out=out)
    torch.cuda.synchronize()
    graphs = []
    for k in range(4):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            m.mesh(vox[k], dn, mode, out=out)
        graphs.append(g)
    for k in range(8): graphs[k % 4].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(200): graphs[k % 4].replay()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us per batch (graph replay)")
