#!/bin/bash
# Strong-scaling sweep on one box with the driver's arguments: bash scripts/scaling_run.sh <workload> <N...>
# (run under gpurun --gpus 8)
W=${1:-cfg4}; shift
STEPS=${STEPS:-20}; WARM=${WARM:-5}
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --workload $W --steps $STEPS --warmup $WARM > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --workload $W --steps $STEPS --warmup $WARM > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
  fi
  echo "rc=$?"
  tail -1 gpurun_out/scale_${W}_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$W N=$N', 'ms/step', round(d['ms_per_step'],2), 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'patch-iters/s %.4e' % d['value'], 'e2e %.4e' % d['e2e']['value'], 'parity', d.get('sharded_parity_rel_l2'), d['clocks'])" || tail -3 gpurun_out/scale_${W}_n$N.err
done
