"""Micro-benchmark of the weight-gradient GEMM (CUDA events, L2 flushed): PASNL_WGRAD_TMA=1 vs 0.
   python tools/bench_wgrad.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(524288, 64, 64), (262144, 128, 128), (262144, 64, 128), (131072, 128, 128), (16384, 2048, 128), (4096, 4096, 256),
          (4096, 512, 1024), (16384, 256, 512), (16384, 128, 256), (16384, 128, 128), (4096, 256, 512), (4096, 256, 256),
          (4096, 64, 256), (16384, 32, 128), (1310720, 32, 32), (4194304, 32, 32)]


def child():
    import torch
    from pointasnl_b200 import _ops
    from tools.bench_op import timeit
    out = {}
    for rows, m, n in SHAPES:
        x = torch.randn(rows, m, device="cuda")
        dy = torch.randn(rows, n, device="cuda")
        med, _ = timeit(lambda: _ops.gemm_wgrad(x, dy), iters=8)
        out[f"{rows}x{m}x{n}"] = med
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = {}
    for tag, env in (("tma", {"PASNL_WGRAD_TMA": "1"}), ("old", {"PASNL_WGRAD_TMA": "0"})):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        if r.returncode != 0:
            print(tag, r.stderr[-1500:])
            continue
        res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    print("shape".ljust(20) + "".join(t.rjust(9) for t in res))
    for k in res.get("old", res.get("tma", {})):
        print(k.ljust(20) + "".join(f"{res[t].get(k, float('nan')):9.1f}" for t in res))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_wgrad.json"), "w"), indent=1)
