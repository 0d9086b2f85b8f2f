#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-d21}
timeout 600 python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 --per-op > gpurun_out/${T}_bench_s.json 2> gpurun_out/${T}_bench_s.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_s.json') if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d.get('ms_per_step_blocks',[])])
PY
