# headline config at 4096 and 65536 envs, device-resident and end to end (one line each)
run() { echo "== $*"; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d['detail']['geometry'])"; }
run A=1
run MJB_BENCH_ENVS=65536
