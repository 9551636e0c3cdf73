run() { echo "== $*"; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d['detail']['geometry'])"; }
for m in 2 4 5 1; do run MJB_LOCKSTEP=$m; run MJB_LOCKSTEP=$m MJB_BENCH_ENVS=65536; done
