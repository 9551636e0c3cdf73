run() { echo "== $*"; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d['detail']['geometry'])"; }
run A=1
run MJB_LOCKSTEP=0
run MJB_LOCKSTEP=3
run MJB_LOCKSTEP=1
run MJB_GROUPS=2
run MJB_CTAS_PER_SM=2
run MJB_WARPS=10
run MJB_WARPS=7
