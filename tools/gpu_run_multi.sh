#!/bin/bash
# bench.py at N GPUs of one box (gpurun --gpus N -- 'bash tools/gpu_run_multi.sh'): N, then the smaller powers of two, then the
# 2-rank correctness test.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-r02b}
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
if [ $N -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "two_ranks or two_engines or devices_list" > gpurun_out/${T}_pytest_2gpu.log 2>&1; echo "2-gpu tests rc=$?"; tail -3 gpurun_out/${T}_pytest_2gpu.log
fi
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --extras "" --no-cpu-baseline --latency-iters 0 > gpurun_out/${T}_bench_yolo11s_b64_${n}gpu.json 2> gpurun_out/${T}_bench_${n}gpu.err; echo "N=$n rc=$?"
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_yolo11s_b64_${n}gpu.json') if l.startswith('{')][-1])
print('N=${n}: value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d['ms_per_step_blocks']],'ranks',[round(x,3) for x in d['ms_per_step_per_rank']])
print('   e2e',round(d['e2e']['value']),'gather',d['gather']['mode'],'no_gather_ms',d['gather'].get('no_gather_ms_per_step'),'clocks',d['clocks']['sm_mhz'],d['clocks']['reasons'])
"
  fi
done
timeout 500 python bench.py --extras "" --no-cpu-baseline --latency-iters 0 > gpurun_out/${T}_bench_1gpu_samebox.json 2> gpurun_out/${T}_bench_1gpu_samebox.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu_samebox.json'))
print('N=1: value',round(d['value']),'ms',round(d['ms_per_step'],3),'blocks',[round(x,3) for x in d['ms_per_step_blocks']],'e2e',round(d['e2e']['value']))
"
