#!/bin/sh
# Usage (on the GPU box): tools/try_depth.sh "DEPTH SLOTS FRAMES SLOTFRAMES" ...
export FID_BENCH_SKIP_CPU=1
for cfg in "$@"; do
  set -- $cfg
  FID_BENCH_DEPTH=$1 FID_SLOTS=$2 FID_BENCH_FRAMES=$3 FID_BENCH_SLOT=$4 timeout 300 python bench.py --steps 8 --warmup 3 2>>gpurun_out/depth_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$cfg', 'fps %.0f e2e %.0f ms/step %.2f e2e ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
done
