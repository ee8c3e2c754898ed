#!/bin/bash
# e2e of library variants: tools/debug/e2e_ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  VXP_LIB=$PWD/voxelphile_b200/$lib python bench.py --steps 100 --no-cpu --no-also --e2e-steps 24 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['e2e']['value']), round(d['e2e']['frac_of_upload_bound'],3))"
  VXP_LIB=$PWD/voxelphile_b200/$lib python bench.py --workload terrain_culled --steps 100 --no-cpu --no-also --e2e-steps 24 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib culled', round(d['e2e']['value']), round(d['e2e']['frac_of_upload_bound'],3))"
done
