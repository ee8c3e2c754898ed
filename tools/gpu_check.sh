#!/bin/bash
# One GPU-box pass used while tuning: parity suite, the greedy/culled bench lines, the per-phase profile.
# usage: tools/gpu_check.sh TAG [pytest -k expression]
tag=$1; sel=${2:-}
if [ -n "$sel" ]; then python -m pytest tests -m gpu -x -q -k "$sel" 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
else python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_tests.log; fi
python bench.py --workload terrain_greedy --steps 400 --no-cpu --e2e-steps 0 > gpurun_out/${tag}_greedy.json 2> gpurun_out/${tag}.err
python tools/phase_prof.py > gpurun_out/${tag}_phase.txt 2>> gpurun_out/${tag}.err
python tools/summarize.py gpurun_out/${tag}_greedy.json >> gpurun_out/${tag}_tests.log 2>&1
