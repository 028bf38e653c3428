#!/bin/bash
# usage: bash tools/gpu_multi_quick.sh N  (under gpurun --gpus N): DDP check + one N-rank bench line, nothing else
N=${1:-2}
mkdir -p gpurun_out
timeout -k 5 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1
echo "ddp_check exit=$?"; grep -E "ddp_check|DDP_CHECK" gpurun_out/ddp_check_$N.log
timeout -k 5 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --sustained-s 0 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "bench $N exit=$?"; grep -h '"metric"' gpurun_out/scale_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), 'clips/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1))"
