#!/bin/bash
# A/B of routing switches on the cls step (one GPU): each line = one bench run with the given environment
mkdir -p gpurun_out
ab() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name rc=$? $(python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/ab_$name.json') if l.startswith('{')][-1]
print(d['ms_per_step'], d['gpu_launches'])")"
}
ab all_on A=1
ab kslabs_off PASNL_GEMM_WS_KSLABS=0
ab smalln_off PASNL_GEMM_WS_SMALLN=0
ab both_off PASNL_GEMM_WS_KSLABS=0 PASNL_GEMM_WS_SMALLN=0
timeout 300 python -m pytest tests/test_gpu_gemm_ws.py -q 2>&1 | tail -2
