cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -x -k "dwconv" 2>&1 | tail -2
python tools/dw_probe.py 2>&1 | grep "^dw"
Y11_LIB=$PWD/yolo_infer_b200/_lib/liby11_dwprobe.so Y11_DW_DBG=2 python tools/dw_probe.py 64 80 80 128 2>&1 | grep "^dw"
Y11_AUTOTUNE=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:stem_kernel -s 2 -c 1 -f -o gpurun_out/stemprobe python bench.py --extras "" --no-cpu-baseline --skip-e2e --latency-iters 0 --steps 2 --warmup 1 --repeats 1 > gpurun_out/stemprobe_ncu.log 2>&1; echo "ncu rc=$?"
