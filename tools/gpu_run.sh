#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-d19}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/r02b_bench_yolo11s_b64.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02b_bench_yolo11s_b64.json') if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d.get('ms_per_step_blocks',[])], 'e2e', round(d['e2e']['value']), round(d['e2e'].get('sync_value',0)))
print({k:(round(v['value']) if isinstance(v,dict) else v) for k,v in d.items() if k.startswith('value_')})
print(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['hbm_side']['frac'], d['clocks'])
print(d['roofline_post'])
print(d['latency_b1']['device_ms_p50'], d['tensor_source']['value'], d['gpu_launches'])
r=json.loads([l for l in open('gpurun_out/r02b_bench_reference.json') if l.startswith('{')][-1]); print('ref', r['value'])
PY
