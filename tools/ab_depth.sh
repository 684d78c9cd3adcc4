for cfg in "4 2" "6 3" "8 4"; do set -- $cfg; FID_SLOTS=$1 FID_BENCH_DEPTH=$2 FID_BENCH_SKIP_CPU=1 timeout 200 python bench.py > gpurun_out/abd_$1_$2.json 2> gpurun_out/abd_$1_$2.err; python -c "
import json;d=json.load(open('gpurun_out/abd_$1_$2.json'));print('slots $1 depth $2',round(d['value']),round(d['e2e']['value']),d['roofline']['launch_ms'])" || tail -3 gpurun_out/abd_$1_$2.err; done
