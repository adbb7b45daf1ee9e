"""Micro-benchmark of the rows GEMM kernels (CUDA events, L2 flushed): the routed entry point with PASNL_GEMM_WS=1/0.
   python tools/bench_gemm.py [--json out]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(262144, 128, 128), (524288, 64, 64), (131072, 128, 128), (4194304, 16, 16), (1310720, 32, 32),
          (262144, 64, 128), (262144, 128, 64), (524288, 32, 64)]


def child():
    import torch
    from pointasnl_b200 import _ops
    from tools.bench_op import timeit
    out = []
    for rows, k, n in SHAPES:
        a = torch.randn(rows, k, device="cuda")
        bt = torch.randn(n, k, device="cuda") / k ** 0.5
        st = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
        med, best = timeit(lambda: _ops.gemm_rows(a, bt, col_stats=st))
        by = 4.0 * rows * (k + n)
        out.append({"shape": [rows, k, n], "median_us": med, "best_us": best, "gbs": by / (med * 1e-6) / 1e9,
                    "ws": os.environ.get("PASNL_GEMM_WS", "1")})
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = []
    for ws in ("1", "0"):
        env = dict(os.environ, PASNL_GEMM_WS=ws)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-2000:])
            continue
        res += json.loads(r.stdout.strip().splitlines()[-1])
    for r in res:
        print(json.dumps(r))
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
