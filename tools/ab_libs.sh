for v in "$@"; do cp ab_libs/$v.so fiducials_b200/libfiducials_b200.so; FID_BENCH_SKIP_CPU=1 timeout 200 python bench.py 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$v',round(d['value']),round(d['e2e']['value']),round(d['roofline']['launch_ms'],4))"; done
