#!/bin/bash
# e2e experiment: the stream-mode e2e leg of bench.py with different host-batch chunk counts (Y11_CHUNKS)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-e2e}
for ck in default 1 2; do
  if [ $ck = default ]; then unset Y11_CHUNKS; else export Y11_CHUNKS=$ck; fi
  timeout 600 python bench.py --extras "" --no-cpu-baseline --latency-iters 0 > gpurun_out/${T}_ck${ck}.json 2> gpurun_out/${T}_ck${ck}.err; echo "bench chunks=$ck rc=$?"
done
python - <<PY
import json
for ck in ("default", "1", "2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/${T}_ck{ck}.json") if l.startswith("{")][-1])
        print(ck, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "sync", round(d["e2e"]["sync_value"]), "steps", d["e2e"]["steps"])
    except Exception as e:
        print(ck, "failed", e)
PY
