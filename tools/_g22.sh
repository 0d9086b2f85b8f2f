timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for m in n s m; do
for c3k in 1 0; do
Y11_C3K_LANES=$c3k timeout 300 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline --skip-e2e --latency-iters 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m c3k_lanes=$c3k', round(d['value']), round(d['ms_per_step'],3))"
done
done
