#!/bin/bash
# scratch driver for one gpurun call: tools/gpu_run.sh <tag> ; edit the body below per call
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rf > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
tail -6 gpurun_out/c8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c8_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c8_smoke.log; tail -2 gpurun_out/c8_smoke.log
timeout 600 python tools/post_scale.py > gpurun_out/r02_post_scale_m1280.json 2> gpurun_out/c8_post.err; cat gpurun_out/r02_post_scale_m1280.json
timeout 600 python tools/post_scale.py --cls-prior 0.00005 > gpurun_out/r02_post_scale_m1280_sparse.json 2>> gpurun_out/c8_post.err; cat gpurun_out/r02_post_scale_m1280_sparse.json
timeout 2400 bash tools/profile_round2.sh r02a > gpurun_out/c8_profile.log 2>&1; tail -5 gpurun_out/c8_profile.log
