#!/bin/bash
# quick A/B of mesh-kernel variants on the config-2/3 batch: prints culled / greedy ms per step per library
for lib in "$@"; do
  VXP_LIB=$PWD/voxelphile_b200/$lib python -c "from voxelphile_b200 import Mesher,_lib; m=Mesher(0); l=_lib.load(); print('$lib', 'smem culled/greedy', l.vxp_debug_smem_bytes(0), l.vxp_debug_smem_bytes(1), 'CTAs/SM stored c/g', l.vxp_debug_occupancy(0,0), l.vxp_debug_occupancy(1,0), 'fused c/g', l.vxp_debug_occupancy(0,1), l.vxp_debug_occupancy(1,1))"
  VXP_LIB=$PWD/voxelphile_b200/$lib timeout -k 10 200 python bench.py --no-cpu --e2e-steps 0 --steps 200 > gpurun_out/qb_$lib.json 2>/dev/null
  python - "$lib" <<'PY'
import json,sys
lib=sys.argv[1]
d=json.loads(open(f"gpurun_out/qb_{lib}.json").read().strip().splitlines()[-1])
a=d["also"]
print(f"{lib:20s} culled {d['ms_per_step']*1e3:7.1f} us frac {d['roofline']['frac']:.3f} | greedy {a['greedy']['ms_per_step']*1e3:7.1f} us frac {a['greedy']['frac']:.3f} | fused_greedy {a['fused_terrain_greedy']['ms_per_step']*1e3:7.1f} us | terrain {a['terrain_fill']['ms_per_step']*1e3:7.1f} us")
PY
done
