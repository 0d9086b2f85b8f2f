timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for m in n s; do
timeout 300 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline --latency-iters 0 --per-op 2> gpurun_out/u21_$m.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
grep -E "dwconv|stem" gpurun_out/u21_$m.err | head -8
done
