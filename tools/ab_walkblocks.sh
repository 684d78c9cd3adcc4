for cfg in "4 4" "4 1" "4 2" "2 1" "2 2"; do set -- $cfg; FID_WALK_BLOCKS_R2=$1 FID_WALK_BLOCKS_R3=$2 FID_BENCH_SKIP_CPU=1 timeout 120 python bench.py --steps 12 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('r2 x$1 r3 x$2',round(d['value']),round(d['e2e']['value']),d['config'].get('markers_found_per_step'))"; done
