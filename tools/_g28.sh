bash tools/profile_round.sh r01g 2>&1 | tail -2
run() { tag=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/r01g_bench_$tag.err | tail -1 > gpurun_out/r01g_bench_$tag.json; python -c "
import json; d=json.loads(open('gpurun_out/r01g_bench_$tag.json').read()); print('$tag', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'conv TF', round(d['roofline']['achieved'],1), 'lat', (d.get('latency_b1') or {}).get('device_ms_p50'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; }
run yolo11n_b64 --per-op
run yolo11s_b64 --model s --per-op
run yolo11m_b64 --model m --no-cpu-baseline --steps 10 --warmup 3
run yolo11m_1280_b16 --model m --imgsz 1280 --batch 16 --no-cpu-baseline --steps 10 --warmup 3
run yolo11x_b1 --model x --batch 1 --no-cpu-baseline --steps 50 --warmup 5
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01g_bench_reference.json 2>/dev/null
