"""Per-item start / end times of the fused terrain+mesh kernel on the 16^3 block (profiling build); GPU box."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
os.environ["VXP_LIB"] = os.path.join(ROOT, "voxelphile_b200", "libvxp_prof.so")
from voxelphile_b200 import Mesher, grid_coords, GREEDY, CULLED
from voxelphile_b200 import _lib
m = Mesher(0); m.use_current_torch_stream(); lib = _lib.load()
if len(sys.argv) > 1:      # rank R of 8 of the 65^3 world (what one GPU of an 8-GPU run meshes)
    wc = grid_coords((-32, -32, -32), (33, 33, 33)); r = int(sys.argv[1]); per = (wc.shape[0] + 7) // 8
    coords = np.ascontiguousarray(wc[r * per:min(wc.shape[0], (r + 1) * per)])
else:
    coords = grid_coords((0, 0, 0), (16, 16, 16))
dc = torch.from_numpy(coords).cuda()
N = coords.shape[0]
out = m.alloc_mesh(N, 3_000_000)
NC = 65536
lib.vxp_debug_read_prof_time.argtypes = [C.c_void_p, C.c_int]
for mode, name in ((GREEDY, "greedy"), (CULLED, "culled")):
    for k in range(3): m.terrain_mesh(dc, 42, mode, out=out)
    torch.cuda.synchronize()
    tb = np.zeros((NC, 2), dtype=np.uint64); lib.vxp_debug_read_prof_time(tb.ctypes.data, NC)   # clears
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.terrain_mesh(dc, 42, mode, out=out); e1.record(); torch.cuda.synchronize()
    lib.vxp_debug_read_prof_time(tb.ctypes.data, NC)
    t0 = tb[:N, 0].astype(np.int64); t1 = tb[:N, 1].astype(np.int64)
    sel = t0 > 0
    base = t0[sel].min(); st = (t0[sel] - base) / 1e3; en = (t1[sel] - base) / 1e3; d = en - st
    q = out.meta[:, 1].cpu().numpy()[sel]
    print(f"== {name}: whole call {e0.elapsed_time(e1)*1e3:.1f} us; items {sel.sum()}; item us mean {d.mean():.2f} p50 {np.median(d):.2f} p99 {np.percentile(d,99):.2f} max {d.max():.2f}")
    print(f"   first start 0, last start {st.max():.1f}, first end {en.min():.1f}, last end {en.max():.1f}; sum of items / 592 = {d.sum()/592:.1f} us")
    edges = np.arange(0, en.max() + 5, 5.0)
    print("   items alive per 5 us bin:", [int(((st < e + 5) & (en > e)).sum()) for e in edges])
    top = np.argsort(-d)[:6]
    idx = np.nonzero(sel)[0]
    for i in top: print(f"   item chunk {idx[i]} coord {coords[idx[i]].tolist()} dur {d[i]:.1f} us start {st[i]:.1f} quads {q[i]}")
    print("   corr(duration, quads) =", np.corrcoef(d, q)[0, 1])
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"fused_items_{name}_{sys.argv[1] if len(sys.argv) > 1 else 'block'}.npz"),
                        coords=coords[idx], dur=d, start=st, quads=q)
