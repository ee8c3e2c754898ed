import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
os.environ["VXP_LIB"] = os.path.join(ROOT, "voxelphile_b200", "libvxp_prof.so")
from voxelphile_b200 import Mesher, _lib
p, palette, mode = float(sys.argv[1]), sys.argv[2], int(sys.argv[3])
PALETTES = {"small": [1, 2, 3, 4], "large": [1000, 1001, 1002, 0xFFFF], "high_bits": [3, 11, 0x103, 0x8003]}
rng = np.random.default_rng(int(p * 1000) + mode)
n = 24 if palette == "small" else 8
ids = np.array(PALETTES[palette], dtype=np.uint16)
vox = np.where(rng.random((n, 32768)) < p, ids[rng.integers(0, 4, (n, 32768))], 0).astype(np.uint16)
nbr = rng.integers(-1, n, size=(n, 6)).astype(np.int32)
m = Mesher(0); lib = _lib.load()
out = m.mesh(torch.from_numpy(vox.view(np.int16)).cuda(), torch.from_numpy(nbr).cuda(), mode, capacity=n * 98304)
torch.cuda.synchronize()
buf = np.zeros((64, 16), dtype=np.uint64)
lib.vxp_debug_read_prof.argtypes = [C.c_void_p, C.c_int]
lib.vxp_debug_read_prof(buf.ctypes.data, 64)
for i in range(n):
    v = int(buf[i, 10]) - 1000
    print(i, "x_pending", v & 1, "with_eq", (v >> 1) & 1, "with_planes", (v >> 2) & 1, "all_solid", (v >> 3) & 1, "single", (v >> 4) & 1, "small", (v >> 5) & 1, "quads", int(out.meta[i, 1]))
