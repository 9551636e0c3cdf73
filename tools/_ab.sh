for E in 4096 65536; do
echo "E=$E $(MJB_BENCH_ENVS=$E timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])')"
done
