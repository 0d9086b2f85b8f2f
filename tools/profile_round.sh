#!/bin/bash
# ncu evidence for one round (run on the GPU box: gpurun -- 'bash tools/profile_round.sh r01e').  Each ncu pass runs only after
# the plain run of the same command exited 0; the tuned launch variants come from a tune cache written by the plain run, so
# ncu sees the same plan without the autotuner's timing launches.
TAG=${1:-rXX}
OUT=gpurun_out
export Y11_TUNE_CACHE=$OUT/${TAG}_tune.json
for m in n s; do
  CMD="python bench.py --model $m --steps 2 --warmup 1 --skip-condition --skip-e2e"
  timeout 300 $CMD --dump-ops $OUT/${TAG}_ops_$m.json > $OUT/${TAG}_plain_$m.log 2>&1 || { echo "plain run failed ($m)"; tail -5 $OUT/${TAG}_plain_$m.log; exit 1; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_yolo11${m}_b64.csv $CMD > $OUT/${TAG}_ncu1_$m.log 2>&1
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -c 81 --csv --log-file $OUT/${TAG}_conv_dram_$m.csv $CMD > $OUT/${TAG}_ncu2_$m.log 2>&1
done
CMD="python bench.py --model s --steps 2 --warmup 1 --skip-condition --skip-e2e"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -c 10 -o $OUT/${TAG}_conv_s_full $CMD > $OUT/${TAG}_ncu3.log 2>&1
ls -la $OUT | grep $TAG
