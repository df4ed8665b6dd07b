#!/bin/bash
# usage: ab_lib.sh "<bench args>" lib1 lib2 ...
args=$1; shift
for lib in "$@"; do
  p=generativedensification_b200/build/variants/$lib.so
  [ "$lib" = default ] && p=generativedensification_b200/libgdr.so
  GDR_LIB=$p python bench.py --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'views/s %.0f' % d['value'], 'batched %.0f' % d['batched']['value'], {k: round(v*1000,1) for k,v in d['stage_ms_per_launch'].items()})
"
done
