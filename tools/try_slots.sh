#!/bin/sh
# Usage (on the GPU box): tools/try_slots.sh "FRAMES SLOT NSLOTS" ...   -- bench each combination
export FID_BENCH_SKIP_CPU=1
for cfg in "$@"; do
  set -- $cfg
  FID_BENCH_FRAMES=$1 FID_BENCH_SLOT=$2 FID_SLOTS=$3 timeout 300 python bench.py --steps 3 --warmup 3 2>>gpurun_out/slots_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); st=d['roofline']['stage_ms_per_batch']
print('$cfg', 'fps %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'stage sum %.1f' % sum(v for k,v in st.items() if not k.startswith('walk_r')))"
done
