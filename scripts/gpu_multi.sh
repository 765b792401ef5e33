#!/bin/bash
# multi-GPU checks: usage gpu_multi.sh N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or accumulators" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -4
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/multi_${N}.json 2> gpurun_out/multi_${N}.err
grep -c "Init COMPLETE\|comm 0x" gpurun_out/multi_${N}.err; grep -o "NVLS[^ ]*\|via NVLS\|P2P/CUMEM\|Channel [0-9]*/[0-9]* :" gpurun_out/multi_${N}.err | sort | uniq -c | head -5
tail -1 gpurun_out/multi_${N}.json | head -c 2500
