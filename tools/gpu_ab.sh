#!/bin/bash
# A/B on one box: full GPU test suite on the product library, then the YOLO11s value loop alternating between a saved baseline
# library (yolo_infer_b200/_lib/liby11_base.so, Y11_LIB) and the product library.  TAG=<prefix of the output files>.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-ab}
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
fi
BASE=$PWD/yolo_infer_b200/_lib/liby11_base.so
for rep in 1 2; do
  for arm in base new; do
    if [ $arm = base ]; then export Y11_LIB=$BASE; else unset Y11_LIB; fi
    timeout 600 python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 ${PEROP:+--per-op} ${BENCH_ARGS} \
      > gpurun_out/${T}_${arm}${rep}.json 2> gpurun_out/${T}_${arm}${rep}.err; echo "bench $arm$rep rc=$?"
  done
done
unset Y11_LIB
python - <<PY
import json
for rep in (1, 2):
    for arm in ("base", "new"):
        try:
            d = json.loads([l for l in open(f"gpurun_out/${T}_{arm}{rep}.json") if l.startswith("{")][-1])
            print(arm, rep, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "blocks", [round(x, 3) for x in d.get("ms_per_step_blocks", [])])
        except Exception as e:
            print(arm, rep, "failed", e)
PY
