#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rf > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_e2e.py -m gpu -q -s -k "init_b or half_a_pixel or raw_head or config1 or image_fixture" 2>&1 | grep -E "yolo11|init \(B\)|config#1|image_small|passed|failed" > gpurun_out/c2_stats.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c2_smoke.log
for sc in n s; do for init in calibrated survey_b; do
  timeout 300 python tools/drift_table.py --scale $sc --init $init --out gpurun_out/c2_drift_${sc}_${init}.md > /dev/null 2> gpurun_out/c2_drift_${sc}_${init}.err
done; done
timeout 600 python bench.py --skip-e2e --per-op --repeats 3 > gpurun_out/c2_bench_s.json 2> gpurun_out/c2_bench_s.err
tail -8 gpurun_out/c2_pytest.log; tail -3 gpurun_out/c2_smoke.log; cat gpurun_out/c2_stats.log; tail -2 gpurun_out/c2_bench_s.err
