#!/bin/bash
# quick check of a kernel change: the named tests, then one base / new pair of the YOLO11s value loop with per-op tables
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-q}
timeout 600 python -m pytest tests -m gpu -q -x -k "${KEXPR:-emit}" 2>&1 | tail -2
for arm in base new; do
  if [ $arm = base ]; then export Y11_LIB=$PWD/yolo_infer_b200/_lib/liby11_base.so; else unset Y11_LIB; fi
  timeout 300 python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 --per-op > gpurun_out/${T}_${arm}.json 2> gpurun_out/${T}_${arm}.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${T}_${arm}.json") if l.startswith("{")][-1])
print("$arm", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "blocks", [round(x, 3) for x in d.get("ms_per_step_blocks", [])])
PY
  grep -hE " model\.23\.cv3\.[012]\.2 " gpurun_out/${T}_${arm}.err
done
