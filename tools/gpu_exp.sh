#!/bin/bash
# one-off experiment runner: the YOLO11s value loop with per-op table under an environment setting, next to the default
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-x}
for arm in off on off2 on2; do
  case $arm in on*) export $EXP_ENV;; *) unset ${EXP_ENV%%=*};; esac
  timeout 600 python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 --per-op > gpurun_out/${T}_${arm}.json 2> gpurun_out/${T}_${arm}.err; echo "bench $arm rc=$?"
done
python - <<PY
import json
for arm in ("off", "on", "off2", "on2"):
    d = json.loads([l for l in open(f"gpurun_out/${T}_{arm}.json") if l.startswith("{")][-1])
    print(arm, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "blocks", [round(x, 3) for x in d.get("ms_per_step_blocks", [])])
PY
grep -h " model.2.cv2 " gpurun_out/${T}_*.err
