#!/bin/bash
# GPU box, 1 GPU: one `ncu --set full` capture per kernel (two launches each, after the fills), exported as raw/source CSV.
# usage: tools/evidence_b.sh TAG [kernel ...]   (kernels: greedy culled fused fill pack unpack lod; default all)
tag=$1; shift; o=gpurun_out
want=" ${*:-greedy culled fused fill pack unpack lod} "
cap() {  # name target regex
  case "$want" in *" $1 "*) ;; *) return;; esac
  python tools/ncu_target.py $2 6 > $o/${tag}_$1.run 2>&1 || return
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 2 -f -o $o/${tag}_$1 python tools/ncu_target.py $2 6 > $o/${tag}_$1.ncu.log 2>&1
  ncu -i $o/${tag}_$1.ncu-rep --page raw --csv > $o/${tag}_$1_raw.csv 2>/dev/null
  ncu -i $o/${tag}_$1.ncu-rep --page source --print-source cuda,sass --csv > $o/${tag}_$1_source.csv 2>/dev/null
  rm -f $o/${tag}_$1.ncu-rep   # the pull-back limit is 64 MiB: keep the CSV exports only
}
cap greedy greedy mesh_kernel_g
cap culled culled mesh_kernel_c
cap fused fused_greedy terrain_mesh_kernel
cap fill fill terrain_fill
cap pack pack ^pack_chunks
cap unpack unpack unpack_chunks
cap lod lod downsample
ls -la $o | grep ${tag}_ > $o/${tag}_files.txt
