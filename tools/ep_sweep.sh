#!/bin/bash
# usage: tools/ep_sweep.sh N "cfg1;cfg2;..."   cfg = "overlap comm_ctas gemm_ctas"
N=$1
IFS=';' read -ra CFGS <<< "${2:-0 0 0;1 148 0;1 256 128;1 512 116}"
port=29560
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  port=$((port+1))
  DCMOE_EP_OVERLAP=$1 DCMOE_EP_COMM_CTAS=$2 DCMOE_EP_GEMM_CTAS=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 12 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('overlap=$1 comm_ctas=$2 gemm_ctas=$3 N=$N ms/step %.3f  %.2f Mtok/s' % (d['ms_per_step'], d['value']/1e6), {k:round(v,2) for k,v in d['stage_ms'].items()})
"
done
