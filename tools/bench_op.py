"""Micro-benchmark of single pasnl entry points (CUDA events, L2 flushed between iterations).
   python tools/bench_op.py nl_fwd|nl_bwd|knn|fps|group [--json out]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pointasnl_b200 import _ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3   # median, best (us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = []
    dev = "cuda"
    if a.op in ("nl_fwd", "nl_bwd"):
        shapes = [(32, 512, 1024, 32), (32, 128, 512, 64), (4, 1280, 10240, 32), (4, 320, 1280, 32), (2, 8192, 65536, 32)]
        for b, s, n, d in shapes:
            q = torch.randn(b, s, d, device=dev, requires_grad=True)
            kv = torch.randn(b, n, 2 * d, device=dev, requires_grad=True)
            sc = d ** -0.5
            if a.op == "nl_fwd":
                with torch.no_grad():
                    med, best = timeit(lambda: _ops.NLCore.apply(q, kv, sc))
                flops = 4.0 * b * s * n * d
            else:
                out = _ops.NLCore.apply(q, kv, sc)
                g = torch.randn_like(out)
                med, best = timeit(lambda: torch.autograd.grad(out, (q, kv), g, retain_graph=True))
                flops = 10.0 * b * s * n * d
            res.append({"op": a.op, "shape": [b, s, n, d], "median_us": med, "best_us": best,
                        "tflops": flops / (med * 1e-6) / 1e12, "impl": os.environ.get("PASNL_NL_FWD", "auto")})
    elif a.op == "knn":
        for b, n, m, k in [(32, 1024, 512, 32), (32, 512, 128, 64), (4, 10240, 10240, 32), (4, 10240, 1280, 32), (2, 65536, 65536, 32)]:
            x = torch.rand(b, n, 3, device=dev)
            qy = x[:, :m].contiguous()
            med, best = timeit(lambda: _ops.knn(k, x, qy), iters=10)
            res.append({"op": "knn", "shape": [b, n, m, k], "median_us": med, "best_us": best,
                        "gdist_per_s": b * n * m / (med * 1e-6) / 1e9})
    elif a.op == "fps":
        for b, n, m in [(32, 1024, 512), (32, 512, 128), (4, 10240, 1280), (4, 1280, 320), (2, 65536, 8192), (2, 8192, 2048)]:
            x = torch.rand(b, n, 3, device=dev)
            med, best = timeit(lambda: _ops.fps(m, x), iters=10)
            res.append({"op": "fps", "shape": [b, n, m], "median_us": med, "best_us": best, "ns_per_round": med * 1e3 / m})
    for r in res:
        print(json.dumps(r))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
