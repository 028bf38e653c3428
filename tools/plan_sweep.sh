#!/bin/bash
# tuning experiment: tile-plan cost-model constants vs measured step time
for l2 in ${L2S:-38 50 64 100}; do for ov in ${OVS:-1.0 0.4 2.0}; do
  r=$(ZNS_PLAN_L2=$l2 ZNS_PLAN_OVH=$ov timeout 120 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3))")
  echo "L2=$l2 OVH=$ov -> $r"
done; done
