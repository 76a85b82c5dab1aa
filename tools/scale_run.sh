#!/bin/bash
# 1 -> N scaling of bench.py on one box (the driver's launch line). Usage: tools/scale_run.sh "1 2 4 8" workload tag
NS=${1:-"1 2"}; W=${2:-reddit_k256}; TAG=${3:-r01}
for n in $NS; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/scale_${TAG}_${W}_n$n.json 2> gpurun_out/scale_${TAG}_${W}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
      bench.py --gpus $n --steps 10 --warmup 3 --workload $W > gpurun_out/scale_${TAG}_${W}_n$n.json 2> gpurun_out/scale_${TAG}_${W}_n$n.err
  fi
  echo "N=$n $(cut -c1-200 gpurun_out/scale_${TAG}_${W}_n$n.json)"; tail -2 gpurun_out/scale_${TAG}_${W}_n$n.err
done
