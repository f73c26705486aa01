#!/bin/bash
# A/B timing of production library variants on one box: tools/ab.sh [bench args] -- lib1 lib2 ...   (empty name = the in-tree libpicstep.so)
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
for lib in "$@"; do
  if [ "$lib" = "main" ]; then unset PICSTEP_LIB; else export PICSTEP_LIB=$PWD/picongpu_b200/variants/libpicstep_$lib.so; fi
  python bench.py --no-e2e --no-cpu --no-parity "${args[@]}" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'ms/step %.2f' % d['ms_per_step'], 'run kernel %.2f ms' % d['roofline']['ms_per_launch'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
