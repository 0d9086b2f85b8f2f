#!/bin/bash
# ncu evidence after the round-2 CUDA-core changes (packed fp32 pairs, depthwise v3, stem write-out): the launch list of a
# YOLO11s pass under load, the per-conv DRAM traffic, and --set full of the non-conv kernels.  The --set full pass over every conv
# (tools/profile_round2b.sh) is not repeated: the conv kernel body only changed in its epilogue arithmetic.  Each ncu pass runs
# after the plain run of the same command exited 0.   gpurun -- 'bash tools/profile_round2c.sh r02c'
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
export Y11_TUNE_CACHE=$OUT/${TAG}_tune.json
COMMON="--steps 1 --warmup 0 --repeats 1 --skip-e2e --no-graph --streams 1 --weights-cache $OUT/${TAG}_weights"
m=s
CMD="python bench.py --model $m $COMMON"
timeout 400 $CMD --dump-ops $OUT/${TAG}_ops_$m.json > $OUT/${TAG}_plain_$m.log 2>&1 || { echo "plain run failed ($m)"; tail -5 $OUT/${TAG}_plain_$m.log; exit 1; }
timeout 400 $CMD > $OUT/${TAG}_plain2_$m.log 2>&1 || { echo "cached plain run failed ($m)"; tail -5 $OUT/${TAG}_plain2_$m.log; exit 1; }
NCONV=$(python -c "import json;print(sum(1 for o in json.load(open('$OUT/${TAG}_ops_$m.json')) if o['kind']=='conv'))")
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_yolo11${m}_b64.csv $CMD > $OUT/${TAG}_ncu1_$m.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:conv_tc_kernel -c $NCONV --csv --log-file $OUT/${TAG}_conv_dram_$m.csv $CMD > $OUT/${TAG}_ncu2_$m.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'decode_list|decode_onepass|sort_nms|stem_kernel|dwconv|attn_kernel|sppf' -c 16 -o $OUT/${TAG}_other_s_full -f $CMD > $OUT/${TAG}_ncu4.log 2>&1
timeout 300 ncu -i $OUT/${TAG}_other_s_full.ncu-rep --page raw --csv > $OUT/${TAG}_other_s_full_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_weights.*.pt $OUT/${TAG}_*_full.ncu-rep
ls -la $OUT | grep $TAG
