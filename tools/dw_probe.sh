cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "dwconv" 2>&1 | tail -2
TAG=r02d bash tools/gpu_run_multi.sh
