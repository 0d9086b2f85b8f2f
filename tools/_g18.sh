export Y11_TUNE_CACHE=gpurun_out/u15_tune.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"decode_onepass" -s 2 -c 1 -o gpurun_out/u18_decode python bench.py --steps 3 --warmup 1 --skip-e2e > gpurun_out/u18a.log 2>&1
export Y11_TUNE_CACHE=gpurun_out/r01e_tune.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"stem_kernel|dwconv_tma|attn_kernel|letterbox" -c 7 -o gpurun_out/u18_misc python bench.py --steps 2 --warmup 1 --skip-e2e --skip-condition > gpurun_out/u18b.log 2>&1
ls -la gpurun_out/u18_*.ncu-rep
