#!/bin/bash
# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N): DDP check on N ranks + scaling bench at 1,2,4,..,N
N=${1:-2}
mkdir -p gpurun_out
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1
echo "ddp_check exit=$?"; grep -E "ddp_check|DDP_CHECK" gpurun_out/ddp_check_$N.log
n=1
while [ $n -le $N ]; do
  if [ $n -eq 1 ]; then
    timeout -k 5 150 python bench.py --gpus 1 --steps 30 --warmup 5 --no-extras > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29512+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-extras > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "bench $n exit=$?"; grep -h '"metric"' gpurun_out/scale_$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), 'clips/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1))"
  n=$((n*2))
done
# same N with the NCCL all-reduce + local Adam path, for comparison
ZNS_P2P_ADAM=0 timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus $N --steps 30 --warmup 5 --no-extras > gpurun_out/scale_${N}_nccl.json 2> gpurun_out/scale_${N}_nccl.err
echo "bench $N (NCCL path) exit=$?"; grep -h '"metric"' gpurun_out/scale_${N}_nccl.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), 'clips/s', round(d['ms_per_step'],3), 'ms/step')"
tail -3 gpurun_out/scale_$N.err
