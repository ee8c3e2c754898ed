"""Does meshing overlap when two contexts mesh different batches on two streams?  (culled = scan + mesh launches)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, CULLED, GREEDY
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
ms = [Mesher(0) for _ in range(2)]
ss = [torch.cuda.Stream() for _ in range(2)]
ms[0].use_current_torch_stream()
vox = [ms[0].terrain_fill(dc, 42 + k) for k in range(4)]
torch.cuda.synchronize()
for m, s in zip(ms, ss): m.use_stream(s)
outs = [m.alloc_mesh(4096, 4_000_000) for m in ms]
for mode, name in ((CULLED, "culled"), (GREEDY, "greedy")):
    for nstreams in (1, 2):
        for k in range(8): ms[k % nstreams].mesh(vox[k % 4], dn, mode, out=outs[k % nstreams])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        for s in ss[:nstreams]: s.wait_event(a)
        N = 200
        for k in range(N): ms[k % nstreams].mesh(vox[k % 4], dn, mode, out=outs[k % nstreams])
        for s in ss[:nstreams]:
            e = torch.cuda.Event(); e.record(s); torch.cuda.current_stream().wait_event(e)
        b.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        print(f"{name} {nstreams} stream(s): {a.elapsed_time(b) / N * 1e3:.1f} us per batch")
