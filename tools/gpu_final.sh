cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02c_bench_yolo11s_b64.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$?"
bash tools/profile_round2c.sh r02c > gpurun_out/r02c_profile.log 2>&1; echo "profile rc=$?"
tail -3 gpurun_out/r02c_profile.log
