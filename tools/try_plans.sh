#!/bin/sh
# Usage (on the GPU box): tools/try_plans.sh plan1 plan2 ...   -- bench each FID_WALK_PLAN, print fps
export FID_BENCH_SKIP_CPU=1
for plan in "$@"; do
  FID_WALK_PLAN="$plan" timeout 300 python bench.py --steps 6 --warmup 3 2>>gpurun_out/plans_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$plan', 'fps %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
