"""Times the secondary kernels of one library build (VXP_LIB): terrain fill, pack, unpack, LOD on the config-2 batch; GPU box."""
import os, sys, hashlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from voxelphile_b200 import Mesher, grid_coords
m = Mesher(0); m.use_current_torch_stream()
coords = grid_coords((0, 0, 0), (16, 16, 16)); dc = torch.from_numpy(coords).cuda()
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
torch.cuda.synchronize()
def timed(fn, steps):
    for s in range(5): fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps): fn(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3
res = {}
res["fill_us"] = timed(lambda s: m.terrain_fill(dc, 100 + s, out=vox[s % 4]), 100)
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
sheet = torch.from_numpy(grid_coords((0, 8, 0), (64, 9, 64))).cuda()     # every chunk its own column
res["fill_sheet_us"] = timed(lambda s: m.terrain_fill(sheet, 100 + s, out=vox[s % 4]), 100)
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
res["fill_hash"] = hashlib.sha1(vox[1].cpu().numpy().tobytes()).hexdigest()[:12]
res["pack_us"] = timed(lambda s: m.pack(vox[s % 4]), 40)
pk, po, pt = m.pack(vox[0]); outs = [torch.empty_like(vox[0]) for _ in range(4)]
res["unpack_us"] = timed(lambda s: m.unpack(pk, po, out=outs[s % 4]), 100)
res["unpack_ok"] = bool(torch.equal(outs[0], vox[0]))
k8 = np.array([[(2 * X + (c & 1)) + 16 * ((2 * Y + ((c >> 1) & 1)) + 16 * (2 * Z + (c >> 2))) for c in range(8)]
               for Z in range(8) for Y in range(8) for X in range(8)], dtype=np.int32)
dk8 = torch.from_numpy(k8).cuda(); lo = [torch.empty((512, 32768), dtype=torch.int16, device="cuda") for _ in range(4)]
res["lod_us"] = timed(lambda s: m.downsample(vox[s % 4], dk8, out=lo[s % 4]), 100)
big = torch.empty(4 * 268435456 // 4, dtype=torch.int32, device="cuda").view(4, -1)
res["memset_268MB_us"] = timed(lambda s: big[s % 4].zero_(), 100)
res["copy_268MB_us"] = timed(lambda s: outs[s % 4].copy_(vox[(s + 1) % 4]), 100)
print(os.path.basename(os.environ.get("VXP_LIB", "libvxp.so")), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()})
