#!/bin/bash
# Final evidence of a tree on one GPU: full GPU test suite, smoke, the default bench line, the reference arm, the ncu passes of
# tools/profile_round2c.sh.   gpurun -- 'TAG=r02d bash tools/gpu_final.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-r02d}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/${T}_bench_yolo11s_b64.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? in ${SECONDS}s"
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$? in ${SECONDS}s"
bash tools/profile_round2c.sh $T > gpurun_out/${T}_profile.log 2>&1; echo "profile rc=$?"
