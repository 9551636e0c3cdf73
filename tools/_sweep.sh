for G in 1 2 3 4 7; do for E in 4096 65536; do
echo "G=$G E=$E $(MJB_GROUPS=$G MJB_BENCH_ENVS=$E timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])')"
done; done
