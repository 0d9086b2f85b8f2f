#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rf > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 600 python -m pytest tests/test_gpu_draw.py tests/test_gpu_decode.py tests/test_gpu_e2e.py -m gpu -q -s -k "draw or decode or raw_head" 2>&1 | grep -E "yolo11|LSB|clipped|detections|passed|failed" > gpurun_out/c3_stats.log
for sc in n s; do for init in calibrated survey_b; do
  timeout 300 python tools/drift_table.py --scale $sc --init $init --out gpurun_out/r02_drift_${sc}_${init}.md > /dev/null 2> gpurun_out/c3_drift_${sc}_${init}.err
done; done
timeout 600 python bench.py --skip-e2e --per-op --repeats 3 > gpurun_out/c3_bench_s.json 2> gpurun_out/c3_bench_s.err
tail -12 gpurun_out/c3_pytest.log; cat gpurun_out/c3_stats.log; tail -2 gpurun_out/c3_bench_s.err
