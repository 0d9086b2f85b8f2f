#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_yolo11s_b64_${n}gpu.json 2> gpurun_out/c11_bench_${n}gpu.err; echo "N=$n rc=$?"
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_yolo11s_b64_${n}gpu.json') if l.startswith('{')][-1])
print('N=${n}: value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d['ms_per_step_blocks']],'ranks',[round(x,3) for x in d['ms_per_step_per_rank']])
print('   e2e',round(d['e2e']['value']),'gather',d['gather']['mode'],'no_gather_ms',d['gather'].get('no_gather_ms_per_step'),'clocks',d['clocks']['sm_mhz'],d['clocks']['reasons'])
"
  fi
done
timeout 500 python bench.py --extras "" --no-cpu-baseline > gpurun_out/c11_bench_1gpu.json 2> gpurun_out/c11_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/c11_bench_1gpu.json'))
print('N=1: value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d['ms_per_step_blocks']],'e2e',round(d['e2e']['value']))
"
