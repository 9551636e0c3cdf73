run() { echo "== $*"; env "$@" python bench.py --steps 30 --warmup 10 --no-cpu --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],4), d['detail']['geometry'], d['detail']['ncon_dropped'])"; }
T=$PWD/mujoco_rl_environment_wrapper_b200/libmjb_t640.so
run MJB_BENCH_ENVS=65536
run MJB_BENCH_ENVS=65536 MJB_MAXCON=6
run MJB_BENCH_ENVS=65536 MJB_MAXCON=6 MJB_LIB=$T MJB_WARPS=16
run MJB_BENCH_ENVS=65536 MJB_MAXCON=6 MJB_LIB=$T MJB_WARPS=18
run MJB_BENCH_ENVS=65536 MJB_MAXCON=6 MJB_LIB=$T MJB_WARPS=20
run MJB_BENCH_ENVS=65536 MJB_MAXCON=4 MJB_LIB=$T MJB_WARPS=20
