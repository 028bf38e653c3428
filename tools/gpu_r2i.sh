#!/bin/bash
# 2-GPU session: discriminating DDP check (overlapped exchange default, then ZNS_DP_OVERLAP=0), bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L
P=29541
run_ddp() { timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/ddp_check.py; }
echo "== ddp_check (overlap on)"; run_ddp $P 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -12 | tee gpurun_out/r2i_ddp_overlap.txt
echo "== ddp_check (ZNS_DP_OVERLAP=0)"; ZNS_DP_OVERLAP=0 run_ddp $((P+1)) 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -12 | tee gpurun_out/r2i_ddp_nooverlap.txt
echo "== bench N=1"; timeout 600 python bench.py --no-extras --sustained-s 0 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
for ov in 1 0; do
  echo "== bench N=2 overlap=$ov"
  ZNS_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+5+ov)) bench.py --gpus 2 --no-extras --sustained-s 0 --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
