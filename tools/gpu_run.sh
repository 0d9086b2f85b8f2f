#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-d9}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
tail -c 600 gpurun_out/${T}_bench_ref.json
