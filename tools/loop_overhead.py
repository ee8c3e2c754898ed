"""Why is bench.py's step loop slower than the sum of its kernels?  Times the culled launch loop
(a) bare, (b) with per-step events, (c) with the NVML sampler thread running."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, CULLED
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
m = Mesher(0); m.use_current_torch_stream()
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
out = m.alloc_mesh(4096, 4_000_000)
def loop(steps, events):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if events else None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for s in range(steps):
        if events: ev[s][0].record()
        m.mesh(vox[s % 4], dn, CULLED, out=out)
        if events: ev[s][1].record()
    b.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    k = sum(x.elapsed_time(y) for x, y in ev) / steps * 1e3 if events else float("nan")
    return a.elapsed_time(b) / steps * 1e3, t_host / steps * 1e6, k
for _ in range(50): m.mesh(vox[0], dn, CULLED, out=out)
torch.cuda.synchronize()
for name, steps, events, sampler in [("bare", 2000, False, False), ("events", 2000, True, False), ("events+nvml", 2000, True, True),
                                      ("events", 200, True, False), ("events+nvml", 200, True, True)]:
    sp = None
    if sampler:
        sp = bench.ClockSampler(0); sp.start()
    t0 = time.time()
    gpu, host, kern = loop(steps, events)
    if sp: print("   sampler:", sp.stop(t0, time.time(), "x"))
    print(f"{name:12s} steps {steps:5d}: device {gpu:7.1f} us/step, host issue {host:6.1f} us/step, mean kernel bracket {kern:7.1f} us")
