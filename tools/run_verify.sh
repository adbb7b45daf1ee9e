#!/bin/bash
# one-GPU verification pass: GPU tests, smoke, the four bench workloads (full JSON lines under gpurun_out/)
mkdir -p gpurun_out
tag=${1:-verify}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_pytest.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/${tag}_smoke.log)"
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_${name}_n1.json 2> gpurun_out/${tag}_${name}_n1.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/${tag}_${name}_n1.json') if l.startswith('{')][-1]
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'launches', d.get('gpu_launches'))
except Exception as e:
    print('no line', e)
PY
)"
}
run cls "${@:2}"
run clsAS --AS --no-cpu-baseline
run seg --workload seg --no-cpu-baseline
run kitti --workload kitti --no-cpu-baseline
