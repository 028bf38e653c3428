#!/bin/bash
# 8-GPU session: discriminating DDP check with the overlapped exchange, bench at N=8 with and without the overlap
mkdir -p gpurun_out
N=${1:-8}
P=29561
echo "== ddp_check N=$N (overlap on)"
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/ddp_check.py 2>&1 | grep "ddp_check\|DDP_CHECK\|step-1\|Error\|error" | tee gpurun_out/r2j_ddp_${N}.txt
for ov in 1 0 1 0; do
  echo "== bench N=$N overlap=$ov"
  ZNS_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+5+ov)) bench.py --gpus $N --no-extras --sustained-s 0 --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])" | tee -a gpurun_out/r2j_bench_${N}.txt
done
echo "== bench N=1"; timeout 600 python bench.py --no-extras --sustained-s 0 --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'])" | tee -a gpurun_out/r2j_bench_${N}.txt
