#!/bin/bash
# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1
echo "ddp_check exit=$?"; tail -3 gpurun_out/ddp_check_$N.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout -k 10 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 --no-extras > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "bench $n exit=$?"; grep -h '"metric"' gpurun_out/scale_$n.json | cut -c1-400
done
grep -h -i -E "NVLS|via P2P|NET/" gpurun_out/scale_$N.err | head -5
