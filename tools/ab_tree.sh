for r in 1; do
for d in tmp_ab .; do (cd $d; FID_BENCH_SKIP_CPU=1 timeout 200 python bench.py 2> /dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$d',round(d['value']),round(d['e2e']['value']),d['roofline']['launch_ms'])"); done; done
