#!/bin/bash
# bench.py at N GPUs with the peer-memory ghost refresh and with NCCL send/recv (development aid)
N=$1; shift
for ph in 1 2 0; do
  RXG_PEER_ALLREDUCE=$((ph==1)) RXG_PEER_HALO=$((ph>0)) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+ph)) bench.py --gpus $N --steps 10 --warmup 3 --no-e2e "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('peer=$ph', 'value', round(d['value']/1e6,2), 'M  ms/step', round(d['ms_per_step'],2), d['phase_ms_per_step'], 'cg', d['config']['cg_iterations_per_step'], d['config']['ghost_refresh'], '|', d['config']['cg_allreduce'])"
done
