export Y11_TUNE_CACHE=gpurun_out/u14_tune.json
timeout 300 python bench.py --steps 3 --warmup 1 --skip-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode|sort_nms|scan_chunks" -c 12 --csv --log-file gpurun_out/u14_post.csv python bench.py --steps 3 --warmup 1 --skip-e2e > /dev/null 2>&1
grep -E "decode|sort_nms|scan" gpurun_out/u14_post.csv | awk -F'","' '{print $5, $NF}' | head -12
unset Y11_TUNE_CACHE
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-iters 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n', round(d['value']), 'e2e', round(d['e2e']['value']))"
python tools/e2e_breakdown.py 2>&1 | tail -4
