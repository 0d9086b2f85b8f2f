timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u17_tests.txt 2>&1; tail -5 gpurun_out/u17_tests.txt
export Y11_TUNE_CACHE=gpurun_out/u15_tune.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode|sort_nms|scan_chunks" -c 6 --csv --log-file gpurun_out/u17_post.csv python bench.py --steps 3 --warmup 1 --skip-e2e > /dev/null 2>&1
grep -E "decode|sort_nms|scan" gpurun_out/u17_post.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -6
unset Y11_TUNE_CACHE
for m in n s; do
timeout 300 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'lat', d['latency_b1']['device_ms_p50'])"
done
python tools/e2e_breakdown.py 2>&1 | tail -7
