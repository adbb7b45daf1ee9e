"""Tile-width sweep of the one-tile rows GEMM on the layers with few row tiles, against cuBLAS fp32 (torch.mm).
   python tools/sweep_small_gemm.py   (each NT runs in a child process: PASNL_ROWS_NT is read per call)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(4096, 259, 256), (4096, 256, 256), (4096, 256, 259), (4096, 512, 256), (4096, 4096, 256), (4096, 256, 4096),
          (4096, 131, 64), (4096, 64, 131), (4096, 134, 64), (4096, 131, 128), (4096, 256, 512), (4096, 1024, 512),
          (4096, 512, 1024), (16384, 131, 128), (16384, 128, 131), (16384, 128, 128), (16384, 32, 128), (16384, 128, 32),
          (16384, 6, 32), (16384, 9, 128), (32768, 3, 64), (16384, 256, 128), (16384, 128, 256), (16384, 1024, 128),
          (16384, 128, 2048), (1280, 256, 256), (640, 512, 512), (160, 512, 512)]


def child():
    import torch
    from pointasnl_b200 import _ops
    from tools.bench_op import timeit
    out = {}
    for rows, k, n in SHAPES:
        a = torch.randn(rows, k, device="cuda")
        bt = torch.randn(n, k, device="cuda")
        if os.environ.get("SWEEP_IMPL") == "cublas":
            med, _ = timeit(lambda: torch.mm(a, bt.t()), iters=8)
        else:
            med, _ = timeit(lambda: _ops.gemm_rows(a, bt), iters=8)
        out[f"{rows}x{k}x{n}"] = med
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = {}
    for tag, env in (("cublas", {"SWEEP_IMPL": "cublas"}), ("auto", {}), ("tma_off", {"PASNL_GEMM_TMA": "0"})):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        if r.returncode != 0:
            print(tag, r.stderr[-1500:])
            continue
        res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    keys = list(res.get("cublas", {}))
    print("shape".ljust(20) + "".join(t.rjust(9) for t in res))
    for k in keys:
        print(k.ljust(20) + "".join(f"{res[t].get(k, float('nan')):9.1f}" for t in res))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_small_gemm.json"), "w"), indent=1)
