#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c4_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/c4_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_draw.py tests/test_gpu_decode.py -m gpu -q -s --timeout 600 -rf > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c4_bench_2gpu.json 2> gpurun_out/c4_bench_2gpu.err; echo "rc=$?" >> gpurun_out/c4_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --gather nccl > gpurun_out/c4_bench_2gpu_nccl.json 2> gpurun_out/c4_bench_2gpu_nccl.err; echo "rc=$?" >> gpurun_out/c4_bench_2gpu_nccl.err
timeout 600 python bench.py --extras "" --no-cpu-baseline > gpurun_out/c4_bench_1gpu.json 2> gpurun_out/c4_bench_1gpu.err; echo "rc=$?" >> gpurun_out/c4_bench_1gpu.err
grep -E "passed|failed|FAILED|gather mode|val:" gpurun_out/c4_pytest.log | tail -20
tail -3 gpurun_out/c4_bench_2gpu.err; tail -3 gpurun_out/c4_bench_2gpu_nccl.err
