#!/bin/bash
# Time the step kernels for each tuning variant library (and with / without candidate lists).
#   usage: tools/gpu_variants.sh <tag> <n_col> <variant> [<variant> ...]   ("main" = libtitgpu.so; suffix +L = lists on)
tag=$1; ncol=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/${tag}_variants.jsonl
: > $out
for v in "$@"; do
  lists=0; name=$v
  if [[ "$v" == *+L ]]; then lists=1; name=${v%+L}; fi
  lib=titsolver_b200/libtitgpu.so
  [[ "$name" != main ]] && lib=titsolver_b200/libtitgpu_${name}.so
  echo "== $v" >> gpurun_out/${tag}_variants.err
  TIT_LISTS=$lists TITGPU_LIB=$PWD/$lib timeout 300 python tools/variant_times.py 3 $ncol 3 3 >> $out 2>> gpurun_out/${tag}_variants.err
done
cat $out
