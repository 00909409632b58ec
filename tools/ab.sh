#!/bin/bash
# tools/ab.sh <variant names...> — on the GPU box: parity smoke + headline bench for the default build and each variant
cd "$(dirname "$0")/.."
for v in default "$@"; do
  if [ "$v" = default ]; then unset SOUNDKIT_CUDA_LIB; else export SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_$v.so; fi
  echo "== $v"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python bench.py --steps 10 --warmup 3 --no-extras --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('value', round(d['value']), 'Msamples/s  frac', round(d['roofline']['frac'], 4), ' e2e', round(d['e2e']['value']), 'bit-exact', d['parity']['bit_exact'], d['parity']['e2e_equals_device_path'])
"
done
