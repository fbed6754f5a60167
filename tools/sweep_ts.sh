#!/bin/bash
# rank-walk -> text-walk switch thresholds (zb_mf_scan_k): one short bench per (min, mul)
for mn in 8 12 16; do for ml in 4 6 10 16; do
  ZULTRA_CUDA_TS_MIN=$mn ZULTRA_CUDA_TS_MUL=$ml python bench.py --steps 2 --warmup 2 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['kernels_ms_per_step']
print('ts_min=$mn ts_mul=$ml', j['value'], 'MB/s', 'mf_scan', k.get('mf_scan'), 'mf_text', k.get('mf_text'), 'match', j['stages_ms']['match'])"
done; done
