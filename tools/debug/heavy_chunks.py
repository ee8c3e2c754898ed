"""Which chunks of the config-2/3 batches take a greedy fallback (profiling build counter 14), and what they cost."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
os.environ["VXP_LIB"] = os.path.join(ROOT, "voxelphile_b200", "libvxp_prof.so")
from voxelphile_b200 import Mesher, grid_coords, neighbour_table, GREEDY
from voxelphile_b200 import _lib
m = Mesher(0); m.use_current_torch_stream(); lib = _lib.load()
coords = grid_coords((0, 0, 0), (16, 16, 16)); nbr = neighbour_table(coords)
dc = torch.from_numpy(coords).cuda(); dn = torch.from_numpy(nbr).cuda()
lib.vxp_debug_read_prof.argtypes = [C.c_void_p, C.c_int]
NC = 65536
for seed in (42, 43, 44, 45):
    vox = m.terrain_fill(dc, seed)
    out = m.alloc_mesh(4096, 4_000_000)
    m.mesh(vox, dn, GREEDY, out=out); torch.cuda.synchronize()
    buf = np.zeros((NC, 16), dtype=np.uint64); lib.vxp_debug_read_prof(buf.ctypes.data, NC)
    m.mesh(vox, dn, GREEDY, out=out); torch.cuda.synchronize()
    lib.vxp_debug_read_prof(buf.ctypes.data, NC)
    act = buf[:4096].astype(np.float64)
    heavy = np.nonzero(act[:, 14] > 0)[0]
    tot = act[:, :8].sum(1)
    q = out.meta[:, 1].cpu().numpy()
    print(f"seed {seed}: fallback CTAs {heavy.tolist()} cycles {tot[heavy].astype(int).tolist()} quads {q[heavy].tolist() if len(heavy) else []}; max cycles {int(tot.max())} (cta {int(tot.argmax())})")
