#!/bin/bash
# Strong-scaling sweep of one scene on the GPUs of this box: tools/sweep_scene.sh c5 "1 2 4 8" out.jsonl [steps] [warmup]
# (bench.py's JSON line per GPU count, one per line; N>1 through torch.distributed.run as the driver launches it)
scene=$1; ns=$2; out=$3; steps=${4:-3}; warm=${5:-3}
: > "$out"
for n in $ns; do
  if [ "$n" = 1 ]; then
    python bench.py --scene $scene --gpus 1 --steps $steps --warmup $warm --no-cpu-baseline >> "$out" 2>> "$out.err"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --scene $scene --gpus $n --steps $steps --warmup $warm >> "$out" 2>> "$out.err"
  fi
  tail -c 300 "$out"; echo
done
