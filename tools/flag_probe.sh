#!/bin/bash
# timing experiments only: each NBMF_TP_DEBUG_FLAGS value disables part of the tensor pass (results are garbage)
for f in 0 2 4 8 12 16 32 48 14 62; do
  NBMF_TP_DEBUG_FLAGS=$f timeout 120 python bench.py --workload big --precision 3 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); r=d['roofline']; print('flags', $f, 'ms/step', round(d['ms_per_step'],3), 'rows', round(r['pass_rows_ms'],3), 'cols', round(r['pass_cols_ms'],3))
except Exception as e:
    print('flags', $f, 'failed', l[-300:])
"
done
