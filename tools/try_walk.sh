#!/bin/bash
# sweep of the border-walk plan (rounds / persistence / refill threshold / pass length); prints frames/s and the walk rounds' in-pipeline ms
for cfg in "8,64,512,p0|16|16" "8,p64,p512,p0|16|16" "8,p64,p512,p0|8|16" "8,p64,p512,p0|4|8" "8,p64,p0|8|16" "8,p0|8|16" "4,p32,p256,p0|8|8" "8,p64,p512,p0|12|8"; do
  IFS='|' read plan refill pass <<< "$cfg"
  FID_WALK_PLAN=$plan FID_WALK_REFILL=$refill FID_WALK_PASS=$pass FID_BENCH_SKIP_CPU=1 FID_BENCH_PARITY_FRAMES=0 timeout 120 python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_batch']
print('$cfg', 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'walk', round(s['walk'],2), [round(s['walk_r%d'%i],2) for i in range(4)])"
done
