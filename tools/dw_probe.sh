cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 300 python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 $ARGS > gpurun_out/x2_$name.json 2> gpurun_out/x2_$name.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/x2_$name.json") if l.startswith("{")][-1])
print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "blocks", [round(x, 3) for x in d.get("ms_per_step_blocks", [])])
PY
}
ARGS="" run default A=1
ARGS="" run stemrows4 Y11_STEM_ROWS=4
ARGS="--streams 3" run streams3 A=1
ARGS="" run default2 A=1
ARGS="" run stemrows2 Y11_STEM_ROWS=2
