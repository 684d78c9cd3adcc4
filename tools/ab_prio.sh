for cfg in "0 4" "1 4" "1 6" "1 8" "0 8"; do set -- $cfg; FID_PRIO=$1 FID_SLOTS=$2 FID_BENCH_SKIP_CPU=1 timeout 200 python bench.py > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab_$1_$2.err; python -c "
import json;d=json.load(open('gpurun_out/ab_$1_$2.json'));print('prio $1 slots $2',round(d['value']),round(d['e2e']['value']),d['config'].get('markers_found_per_step'))"; done
