#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp8.py tests/test_gpu_kernels.py -m gpu -q -s --timeout 600 -rf -k "fp8 or progressive or multi_label or nms" > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
grep -E "passed|failed|FAILED|FP8:|Error|error" gpurun_out/c9_pytest.log | head -30
timeout 600 python tools/post_scale.py > gpurun_out/r02_post_scale_m1280_topk.json 2> gpurun_out/c9_post.err; cat gpurun_out/r02_post_scale_m1280_topk.json; tail -3 gpurun_out/c9_post.err
timeout 600 python tools/post_scale.py --cls-prior 0.00005 > gpurun_out/r02_post_scale_m1280_sparse_topk.json 2>> gpurun_out/c9_post.err; cat gpurun_out/r02_post_scale_m1280_sparse_topk.json
