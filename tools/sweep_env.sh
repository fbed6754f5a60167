#!/bin/bash
# usage: sweep_env.sh VAR "v1 v2 ..." [bench args]: one short bench.py run per value of an environment knob, one summary line each
var=$1; vals=$2; shift 2
for v in $vals; do
  env $var=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['kernels_ms_per_step']
print('$var=$v', j['value'], 'MB/s', 'parse_dp', k.get('parse_dp'), 'repair', k.get('parse_repair'), 'verify', k.get('parse_verify'), 'redo', j['counters']['parse_redo'], 'stages', j['stages_ms'])"
done
