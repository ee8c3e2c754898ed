#!/bin/bash
# GPU box with N GPUs: the bench lines of the scaling table (one rank per GPU, torchrun) and the sharded parity test.
# usage: tools/scale.sh TAG N
tag=$1; n=$2; o=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $n "$@"; }
run > $o/${tag}_scale_greedy_${n}gpu.json 2> $o/${tag}_scale_greedy_${n}gpu.err
run --workload terrain_culled --no-cpu > $o/${tag}_scale_culled_${n}gpu.json 2> $o/${tag}_scale_culled_${n}gpu.err
run --workload world_fused --no-cpu > $o/${tag}_scale_world_fused_${n}gpu.json 2> $o/${tag}_scale_world_fused_${n}gpu.err
run --impl reference --steps 3 --warmup 1 > $o/${tag}_scale_reference_${n}gpu.json 2> $o/${tag}_scale_reference_${n}gpu.err
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4 > $o/${tag}_multi_gpu_tests_${n}gpu.log
for f in greedy culled world_fused; do python tools/summarize.py $o/${tag}_scale_${f}_${n}gpu.json >> $o/${tag}_multi_gpu_tests_${n}gpu.log 2>&1; done
