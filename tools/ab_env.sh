#!/bin/bash
# usage: tools/ab_env.sh VAR v1 v2 ...  -> one short bench per value of the environment variable, prints views/s + stage times
var=$1; shift
for val in "$@"; do
  env $var=$val python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$val', 'views/s %.0f' % d['value'], 'batched %.0f' % d['batched']['value'], {k: round(v*1000,1) for k,v in d['stage_ms_per_launch'].items()})
"
done
