python -m pytest tests -m gpu -x -q > gpurun_out/u3_tests.txt 2>&1; tail -5 gpurun_out/u3_tests.txt
for m in n s; do
python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline --latency-iters 0 --per-op 2> gpurun_out/u3_${m}.err | tail -1 > gpurun_out/u3_${m}.json; python -c "import sys,json; d=json.loads(open('gpurun_out/u3_${m}.json').read()); print('$m tuned', d.get('value'), d.get('ms_per_step'), d['roofline']['achieved'], d['e2e']['value'])"
Y11_AUTOTUNE=0 python bench.py --model $m --steps 20 --warmup 5 --skip-e2e --no-cpu-baseline --latency-iters 0 2> /dev/null | tail -1
done
