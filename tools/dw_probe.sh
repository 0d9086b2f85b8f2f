cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for mw in 32 18; do
echo "Y11_DW_TMA_MINW=$mw"
Y11_DW_TMA_MINW=$mw python tools/dw_probe.py 64 20 20 512 64 20 20 128 64 20 20 256 64 40 40 128 2>&1 | grep "^dw"
done
Y11_DW_TMA_MINW=18 python -m pytest tests/test_gpu_kernels.py -q -x -k "dwconv" 2>&1 | tail -2
