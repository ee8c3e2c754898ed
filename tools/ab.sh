#!/bin/bash
# A/B of library variants on config 2 / 3: tools/ab.sh TAG lib1.so lib2.so ...   (GPU box)
tag=$1; shift
for lib in "$@"; do
  VXP_LIB=$PWD/voxelphile_b200/$lib python bench.py --workload terrain_greedy --steps 300 --no-cpu --e2e-steps 0 --no-also > gpurun_out/${tag}_${lib%.so}_g.json 2>> gpurun_out/${tag}.err
  VXP_LIB=$PWD/voxelphile_b200/$lib python bench.py --workload terrain_culled --steps 300 --no-cpu --e2e-steps 0 --no-also > gpurun_out/${tag}_${lib%.so}_c.json 2>> gpurun_out/${tag}.err
  python tools/summarize.py gpurun_out/${tag}_${lib%.so}_g.json gpurun_out/${tag}_${lib%.so}_c.json | grep -v clocks
done > gpurun_out/${tag}_ab.txt 2>&1
