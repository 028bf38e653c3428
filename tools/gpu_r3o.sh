#!/bin/bash
# the driver's multi-GPU invocation, unabridged (default extras on rank 0, sustained leg on every rank), N = 2
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3o_scale2.json 2> gpurun_out/r3o_scale2.err
echo "bench 2 exit=$?"; grep -h '"metric"' gpurun_out/r3o_scale2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), 'clips/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1), 'sustained', round(d['sustained']['value'],1), 'keys', sorted(d.keys()))"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
