timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u19_tests.txt 2>&1; tail -5 gpurun_out/u19_tests.txt
for m in n s; do
timeout 300 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline --per-op 2> gpurun_out/u19_$m.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'lat', d['latency_b1']['device_ms_p50'])"
grep attention gpurun_out/u19_$m.err
done
python tools/e2e_breakdown.py 2>&1 | tail -3
