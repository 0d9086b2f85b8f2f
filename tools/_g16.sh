export Y11_TUNE_CACHE=gpurun_out/u15_tune.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"decode_onepass|sort_nms" -s 2 -c 2 -o gpurun_out/u16_post python bench.py --steps 3 --warmup 1 --skip-e2e > gpurun_out/u16.log 2>&1
ls -la gpurun_out/u16_post.ncu-rep
