python -m pytest tests -m gpu -x -q > gpurun_out/u2_tests.txt 2>&1; tail -5 gpurun_out/u2_tests.txt
for m in n s; do
python bench.py --model $m --steps 20 --warmup 5 --skip-e2e --no-cpu-baseline --latency-iters 0 --per-op 2> gpurun_out/u2_${m}.err | tail -1 > gpurun_out/u2_${m}.json; python -c "import sys,json; d=json.loads(open('gpurun_out/u2_${m}.json').read()); print('$m fold', d.get('value'), d.get('ms_per_step'), d['roofline']['achieved'])"
Y11_FOLD_UP=0 python bench.py --model $m --steps 20 --warmup 5 --skip-e2e --no-cpu-baseline --latency-iters 0 2> /dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m nofold', d.get('value'), d.get('ms_per_step'), d['roofline']['achieved'])"
done
