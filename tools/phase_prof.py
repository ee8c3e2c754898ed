"""Per-phase cycle breakdown of the mesh kernel (profiling build libvxp_prof.so, `make -C voxelphile_b200/csrc prof`).
GPU box only; prints one table per mode for the config-2/3 terrain batch."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["VXP_LIB"] = os.path.join(ROOT, "voxelphile_b200", os.environ.get("VXP_PROF_LIB", "libvxp_prof.so"))
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, CULLED, GREEDY
from voxelphile_b200 import _lib
m = Mesher(0); m.use_current_torch_stream()
lib = _lib.load()
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
vox = [m.terrain_fill(dc, 42 + k) for k in range(4)]
phases = ["convert", "flags+publish", "halo", "equality", "masks", "sweep/count", "stage", "flush"]
counts = {8: "n_air", 9: "n_buried", 10: "n_nofaces", 11: "n_emit", 12: "xface_summary(warp0)", 13: "xface_gather(warp0)", 14: "greedy_heavy_path", 15: "xface_polls(warp0)"}
NC = 8192
lib.vxp_debug_read_prof.argtypes = [C.c_void_p, C.c_int]
for mode, mn in ((CULLED, "culled"), (GREEDY, "greedy")):
    out = m.alloc_mesh(4096, 4_000_000)
    for k in range(3): m.mesh(vox[k % 4], dn, mode, out=out)
    torch.cuda.synchronize()
    buf = np.zeros((NC, 16), dtype=np.uint64)
    lib.vxp_debug_read_prof(buf.ctypes.data, NC)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.mesh(vox[3], dn, mode, out=out); e1.record(); torch.cuda.synchronize()
    lib.vxp_debug_read_prof(buf.ctypes.data, NC)
    act = buf[:4096].astype(np.float64)
    tot = act[:, :8].sum(1)
    kinds = {"air": act[:, 8] > 0, "buried": act[:, 9] > 0, "emit": act[:, 11] > 0}
    print(f"== {mn}: kernel {e0.elapsed_time(e1)*1e3:.1f} us; per-CTA cycles mean {tot.mean():.0f} max {tot.max():.0f}")
    for kn, sel in kinds.items():
        if sel.sum() == 0: continue
        t = tot[sel]
        print(f"   [{kn:6s}] {int(sel.sum()):5d} chunks, cycles mean {t.mean():8.0f} p50 {np.median(t):8.0f} p99 {np.percentile(t,99):8.0f} max {t.max():8.0f}")
        for i, nm in enumerate(phases):
            col = act[sel, i]
            if col.sum() > 0: print(f"        {nm:14s} mean {col.mean():8.0f}  max {col.max():8.0f}")
    for i, nm in counts.items(): print(f"   {nm:20s} {int(act[:, i].sum())}")
    tb = np.zeros((NC, 2), dtype=np.uint64)
    lib.vxp_debug_read_prof_time.argtypes = [C.c_void_p, C.c_int]; lib.vxp_debug_read_prof_time(tb.ctypes.data, NC)
    t0 = tb[:4096, 0].astype(np.int64); t1 = tb[:4096, 1].astype(np.int64); base = t0.min()
    st, en = (t0 - base) / 1e3, (t1 - base) / 1e3
    edges = np.arange(0, en.max() + 5, 5.0)
    conc = [int(((st < e + 5) & (en > e)).sum()) for e in edges]
    print("   CTAs alive per 5 us bin:", conc)
    print(f"   first CTA end {en.min():.1f} us, last CTA start {st.max():.1f} us, last end {en.max():.1f} us; CTA life mean {np.mean(en-st):.2f} us; sum of lives / 444 = {np.sum(en-st)/444:.1f} us")
    if mode == GREEDY:
        wb = np.zeros((NC, 8), dtype=np.uint32)
        lib.vxp_debug_read_prof_warp.argtypes = [C.c_void_p, C.c_int]; lib.vxp_debug_read_prof_warp(wb.ctypes.data, NC)
        w = wb[:4096][kinds["emit"]].astype(np.float64)
        print("   sweep cycles per warp (f = warp), emit chunks: mean", np.round(w.mean(0)).astype(int).tolist(), " mean of max", int(w.max(1).mean()))
