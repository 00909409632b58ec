#!/bin/bash
# tools/ab_gen.sh [variant names...] — general-ratio resampler microbench (tools/kbench_gen.py) for the default build and variants
cd "$(dirname "$0")/.."
for v in default "$@"; do
  if [ "$v" = default ]; then unset SOUNDKIT_CUDA_LIB; else export SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_$v.so; fi
  echo "== $v"; python tools/kbench_gen.py ${KB_SIZES:-16 64 256} | grep -v "^[{}]" | paste - - - - | awk '{print $1, $2, $5}'
done
