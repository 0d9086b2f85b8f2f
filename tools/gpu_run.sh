#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-d15}
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_round2.py -x -q -m gpu -k "every_conv or raw_head or init_b or half_a_pixel" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
for f in 1 0; do
Y11_S2D_COMPACT=$f timeout 600 python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 --per-op > gpurun_out/${T}_bench_s_c$f.json 2> gpurun_out/${T}_bench_s_c$f.err; echo "bench compact=$f rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_s_c$f.json') if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d.get('ms_per_step_blocks',[])])
PY
grep "model.1 \|model.0 " gpurun_out/${T}_bench_s_c$f.err
done
