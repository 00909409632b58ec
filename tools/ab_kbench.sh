#!/bin/bash
# tools/ab_kbench.sh "<kbench groups>" <variant names...> — secondary-kernel microbench for the default build and variants
cd "$(dirname "$0")/.."
groups=$1; shift
for v in default "$@"; do
  if [ "$v" = default ]; then unset SOUNDKIT_CUDA_LIB; else export SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_$v.so; fi
  echo "== $v"
  python tools/kbench.py $groups 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, v in d.items():
    if isinstance(v, dict): print('  %-42s %s' % (k, '  '.join('%s=%.4g' % (a, b) for a, b in v.items() if isinstance(b, (int, float)))))
"
done
