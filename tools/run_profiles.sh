#!/bin/bash
# refresh the per-round profile artefacts of HEAD (one GPU): GPU tests, per-launch event tables of the four workloads,
# the serialised launch list of one cls step under ncu
mkdir -p gpurun_out
tag=${1:-prof}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_pytest.log)"
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-report gpurun_out/${tag}_kernels_events_${name}.json "$@" > gpurun_out/${tag}_${name}_n1.json 2> gpurun_out/${tag}_${name}_n1.err
  echo "$name rc=$? $(cut -c1-200 gpurun_out/${tag}_${name}_n1.json | tail -1)"
}
run cls
run clsAS --AS
run seg --workload seg
run kitti --workload kitti
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches_cls_step.csv python tools/one_step.py --steps 1 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$? $(wc -l < gpurun_out/${tag}_launches_cls_step.csv) lines"
