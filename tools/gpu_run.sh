#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-d8}
timeout 1500 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "class_emit" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
