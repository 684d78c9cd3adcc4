#!/bin/sh
# Usage (on the GPU box): tools/try_plans.sh plan1 plan2 ...   -- bench each FID_WALK_PLAN, print fps + stage times
export FID_BENCH_SKIP_CPU=1
for plan in "$@"; do
  FID_WALK_PLAN="$plan" timeout 300 python bench.py --steps 3 --warmup 3 2>>gpurun_out/plans_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); st=d['roofline']['stage_ms_per_batch']
print('$plan', 'fps %.0f e2e %.0f' % (d['value'], d['e2e']['value']), {k: round(v,2) for k,v in st.items() if v>0.004})"
done
