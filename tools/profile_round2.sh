#!/bin/bash
# ncu evidence for round 2 (run on the GPU box: gpurun -- 'bash tools/profile_round2.sh r02a').  Every ncu pass runs only after the
# plain run of the same command exited 0.  The plain run writes (a) the tuned launch variants (Y11_TUNE_CACHE) and (b) the conditioned
# weights (--weights-cache), so the ncu passes see the same plan and the same ~1300 candidates per image in decode / NMS without the
# autotuner's and the conditioning pass's launches.  --no-graph --streams 1: kernels launched one by one in plan order, so conv launch
# #k of the first pass is conv #k of the op list (tools/ncu_join.py joins them by position).
TAG=${1:-r02a}
OUT=gpurun_out
export Y11_TUNE_CACHE=$OUT/${TAG}_tune.json
COMMON="--steps 1 --warmup 0 --repeats 1 --skip-e2e --no-graph --streams 1 --weights-cache $OUT/${TAG}_weights"
for m in s n m; do
  CMD="python bench.py --model $m $COMMON"
  timeout 400 $CMD --dump-ops $OUT/${TAG}_ops_$m.json > $OUT/${TAG}_plain_$m.log 2>&1 || { echo "plain run failed ($m)"; tail -5 $OUT/${TAG}_plain_$m.log; exit 1; }
  timeout 400 $CMD > $OUT/${TAG}_plain2_$m.log 2>&1 || { echo "cached plain run failed ($m)"; tail -5 $OUT/${TAG}_plain2_$m.log; exit 1; }
done
for m in s n; do
  CMD="python bench.py --model $m $COMMON"
  NCONV=$(python -c "import json;print(sum(1 for o in json.load(open('$OUT/${TAG}_ops_$m.json')) if o['kind']=='conv'))")
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_yolo11${m}_b64.csv $CMD > $OUT/${TAG}_ncu1_$m.log 2>&1
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:conv_tc_kernel -c $NCONV --csv --log-file $OUT/${TAG}_conv_dram_$m.csv $CMD > $OUT/${TAG}_ncu2_$m.log 2>&1
done
# --set full of every conv launch of one YOLO11s step and of one YOLO11m step (tensor-pipe utilisation of the compute-bound layers)
for m in s m; do
  CMD="python bench.py --model $m $COMMON"
  NCONV=$(python -c "import json;print(sum(1 for o in json.load(open('$OUT/${TAG}_ops_$m.json')) if o['kind']=='conv'))")
  timeout 1500 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -c $NCONV -o $OUT/${TAG}_conv_${m}_full -f $CMD > $OUT/${TAG}_ncu3_$m.log 2>&1
  timeout 300 ncu -i $OUT/${TAG}_conv_${m}_full.ncu-rep --page raw --csv > $OUT/${TAG}_conv_${m}_full_raw.csv 2>/dev/null
done
# decode / NMS / letterbox under load: --set full of the post-processing kernels of one YOLO11s step
CMD="python bench.py --model s $COMMON"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'decode_onepass|sort_nms|stem_kernel|dwconv|attn_kernel|sppf' -c 16 -o $OUT/${TAG}_other_s_full -f $CMD > $OUT/${TAG}_ncu4.log 2>&1
timeout 300 ncu -i $OUT/${TAG}_other_s_full.ncu-rep --page raw --csv > $OUT/${TAG}_other_s_full_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_weights.*.pt $OUT/${TAG}_*_full.ncu-rep     # the raw CSV exports stay; gpurun merges at most 64 MiB back
ls -la $OUT | grep $TAG
