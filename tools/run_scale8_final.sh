#!/bin/bash
# final 8-GPU weak-scaling lines of every BASELINE config + strong scaling of config 2 (run under gpurun --gpus 8)
mkdir -p gpurun_out
run() {
  name=$1; g=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/final_${name}_n$g.json 2> gpurun_out/final_${name}_n$g.err
  echo "$name n=$g rc=$? $(cut -c1-170 gpurun_out/final_${name}_n$g.json | tail -1)"
}
run cls 8
run clsAS 8 --AS
run seg 8 --workload seg
run kitti 8 --workload kitti
run cls_strong 8 --scaling strong --batch 32
run cls 2
