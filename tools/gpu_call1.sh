#!/bin/bash
# first GPU call of round 2: suite + stats + smoke + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_e2e.py -m gpu -q -s -k "init_b or half_a_pixel or raw_head or config1 or image_fixture" > gpurun_out/c1_stats.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c1_smoke.log
timeout 900 python bench.py --per-op > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?" >> gpurun_out/c1_bench.err
tail -5 gpurun_out/c1_pytest.log; tail -3 gpurun_out/c1_smoke.log; tail -3 gpurun_out/c1_bench.err
