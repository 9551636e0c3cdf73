for Z in 2 1; do
echo "ZEROCOPY=$Z $(MJB_HOST_ZEROCOPY=$Z timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"])')"
done
