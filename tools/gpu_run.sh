#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-d14}
for st in 1 2; do for f in 1 0; do
Y11_PAIR=$f timeout 600 python bench.py --streams $st --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 > gpurun_out/${T}_s${st}_p$f.json 2> gpurun_out/${T}_s${st}_p$f.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_s${st}_p$f.json') if l.startswith('{')][-1])
print('streams $st pair $f: value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d.get('ms_per_step_blocks',[])])
PY
done; done
for f in 1 0; do
Y11_PAIR=$f timeout 600 python bench.py --model m --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 > gpurun_out/${T}_m_p$f.json 2> gpurun_out/${T}_m_p$f.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_m_p$f.json') if l.startswith('{')][-1])
print('model m pair $f: value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d.get('ms_per_step_blocks',[])])
PY
done
