"""Per-kernel durations of the fused path on a 1/8 slice of the config-5 world (what one of 8 ranks gets).
Run plain for CUDA-event timing of the whole call, or under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from voxelphile_b200 import Mesher, grid_coords, GREEDY
from voxelphile_b200.shard import partition
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = Mesher(0)
wc = grid_coords((-32, -32, -32), (33, 33, 33))
for r in (0, parts // 2):
    lo, hi = partition(wc.shape[0], r, parts)
    c = torch.from_numpy(np.ascontiguousarray(wc[lo:hi])).cuda()
    out = m.alloc_mesh(hi - lo, 4_000_000)
    for _ in range(3): m.terrain_mesh(c, 42, GREEDY, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): m.terrain_mesh(c, 42, GREEDY, out=out)
    b.record(); torch.cuda.synchronize()
    print(f"rank {r}/{parts}: {hi-lo} chunks, {int(out.total.item())} quads, {a.elapsed_time(b)/20*1e3:.1f} us per call, work items {int((out.meta[:,1]>0).sum())}")
