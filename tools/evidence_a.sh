#!/bin/bash
# GPU box, 1 GPU: parity suite, the bench lines kept under profiles/, and the ncu launch list of the default bench command.
# usage: tools/evidence_a.sh TAG
tag=$1; o=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $o/${tag}_tests.log
python bench.py > $o/${tag}_bench_greedy.json 2> $o/${tag}_bench_greedy.err
python bench.py --workload terrain_culled > $o/${tag}_bench_culled.json 2> $o/${tag}_bench_culled.err
python bench.py --workload world_fused > $o/${tag}_bench_world_fused.json 2> $o/${tag}_bench_world_fused.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_reference_arm.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-also --e2e-steps 2 --soak-seconds 0 > $o/${tag}_bench_short.json 2> $o/${tag}_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-also --e2e-steps 2 --soak-seconds 0 > $o/${tag}_ncu_launches.log 2>&1
for f in greedy culled world_fused reference_arm; do python tools/summarize.py $o/${tag}_bench_$f.json >> $o/${tag}_tests.log 2>&1; done
