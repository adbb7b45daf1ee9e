#!/usr/bin/env python
"""bench.py -- point-clouds/s for one training step (forward + loss + backward + gradient all-reduce + Adam) of
pointasnl_cls on ModelNet40-shaped synthetic input, B=32 clouds of N=1024 points per GPU (BASELINE.json configs[1];
--AS gives configs[2]; --workload seg gives configs[3] shape).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU under torchrun
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the box's HOST cores

Prints ONE JSON line (rank 0).  Keys follow the driver contract: value = whole-job clouds/s with inputs resident in
HBM; e2e = the same step fed from pinned host memory with the loss read back every step; roofline = the dominant
pasnl kernel of the step against the measured B200 peaks; cpu_baseline = the CPU restatement timed on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "point-clouds/sec fwd+bwd ModelNet40 B=32 N=1024"
UNIT = "clouds/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cls", choices=["cls", "seg", "kitti"],
                    help="cls = configs[1] (configs[2] with --AS); seg = configs[3] (ScanNet grid shape, B=4 N=10240, rgb); "
                         "kitti = configs[4] (SemanticKITTI grid shape, B=2 N=65536, remission)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch clouds PER GPU; strong: --batch clouds in total, sharded over the ranks "
                         "(per-replica batch-norm statistics: not the single-GPU model at the same global batch)")
    ap.add_argument("--AS", action="store_true", help="adaptive sampling on (BASELINE configs[2])")
    ap.add_argument("--batch", type=int, default=None, help="clouds per GPU (default 32 cls / 4 seg)")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-report", default=None, help="write the per-kernel table (JSON) to this path")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ data
def make_batch(workload, batch, points, seed):
    from tests import synth
    if workload == "cls":
        pc, labels = synth.modelnet_batch(batch, points, seed)
        return pc, labels, None
    if workload == "kitti":
        pts, labels = synth.kitti_batch(batch, points, seed)
        return pts, labels, np.random.default_rng(seed).uniform(1.0, 4.9, labels.shape).astype(np.float32)
    pts, labels, smpw = synth.scannet_batch(batch, points, seed)
    return pts, labels, smpw


# model / data configuration of each workload: TrainStep kwargs, default batch per GPU, default points
WORKLOADS = {
    "cls": dict(model="cls", batch=32, points=1024, kw={}),
    "seg": dict(model="seg", batch=4, points=10240, kw=dict(num_class=21, feature_channel=3, radius=0.07)),
    "kitti": dict(model="seg", batch=2, points=65536, kw=dict(num_class=20, feature_channel=1, radius=0.2)),
}


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ per-kernel model
# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture of the current
# build: profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the raw ncu CSV (keys "entry_point|d0,d1,d2").
def load_ncu_traffic():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    return json.load(open(path)).get("traffic", {})


def gemm_bound(nbytes, flops):
    """Which roof a GEMM launch sits under: 3xTF32 issues 3 tensor passes at half the bf16 rate, so the tensor time is
    6 * flops / bf16_peak; the skinny layers of this path (rows >> k, n) stream their operands and are HBM bound."""
    pk = load_peaks()
    t_tensor = 6.0 * flops / (pk["bf16_tflops"] * 1e12)
    t_hbm = nbytes / (pk["hbm_gbs"] * 1e9)
    return "tensor" if t_tensor > t_hbm else "hbm"


def algorithmic_work(name, a):
    """(bytes, flops, bound) of one launch from its integer arguments (stated in DESIGN.md)."""
    if name == "pasnl_fps":
        b, n, m = a[0:3]
        return 12 * b * n + 4 * b * m, 10 * b * m * n, "latency"
    if name == "pasnl_knn":
        b, n, m, k = a[0:4]
        return 12 * b * (n + m) + 4 * b * m * k, 8 * b * m * n, "fp32"
    if name == "pasnl_ball_query":
        b, n, m, _r, k = a[0:5]
        return 12 * b * (n + m) + 4 * b * m * (k + 1), 9 * b * m * n, "fp32"
    if name in ("pasnl_gather_rows", "pasnl_gather_rows_grad"):
        b, n, c, m = a[0:4]
        return 4 * b * m + 4 * b * n * c + 4 * b * m * c, 0, "hbm"
    if name in ("pasnl_gather_point", "pasnl_gather_point_grad"):
        b, n, m = a[0:3]
        return 4 * b * m + 12 * b * n + 12 * b * m, 0, "hbm"
    if name in ("pasnl_group_point", "pasnl_group_point_grad"):
        b, n, c, m, ns = a[0:5]
        return 4 * b * m * ns + 4 * b * n * c + 4 * b * m * ns * c, 0, "hbm"
    if name == "pasnl_group_concat":
        b, n, c, m, ns = a[0:5]
        w = (3 if a[9] else 0) + 3 + c
        return 4 * b * m * ns + 4 * b * n * (3 + c) + (12 * b * m if a[9] else 0) + 4 * b * m * ns * w, 0, "hbm"
    if name == "pasnl_group_concat_grad":
        b, n, c, m, ns = a[0:5]
        w = (3 if a[8] else 0) + 3 + c
        return 4 * b * m * ns + 4 * b * n * (3 + c) + (12 * b * m if a[8] else 0) + 4 * b * m * ns * w, b * m * ns * w, "hbm"
    if name == "pasnl_three_nn":
        b, n, m = a[0:3]
        return 12 * b * (n + m) + 24 * b * n, 8 * b * n * m, "fp32"
    if name in ("pasnl_three_interpolate", "pasnl_three_interpolate_grad"):
        if name == "pasnl_three_interpolate":
            b, m, c, n = a[0:4]
        else:
            b, n, c, m = a[0:4]
        return 4 * b * m * c + 24 * b * n + 4 * b * n * c, 5 * b * n * c, "hbm"
    if name == "pasnl_as_core_fwd":
        g, k, d = a[0:3]
        return 16 * g * k * d + 4 * g * k * k, 4 * g * k * k * d, "hbm"
    if name == "pasnl_as_core_bwd":
        g, k, d = a[0:3]
        return 28 * g * k * d + 4 * g * k * k, 10 * g * k * k * d, "hbm"
    if name == "pasnl_as_resample_fwd":
        g, k, cg = a[0:3]
        return 4 * g * k * (1 + cg) * 2 + 12 * g * k + 4 * g * k * cg + 4 * g * (3 + cg), 4 * g * k * (1 + cg), "hbm"
    if name == "pasnl_as_resample_bwd":
        g, k, cg = a[0:3]
        return 4 * g * k * (1 + cg) * 2 + 24 * g * k + 8 * g * k * cg + 4 * g * (3 + cg), 6 * g * k * (1 + cg), "hbm"
    if name == "pasnl_nl_core_fwd":
        b, s, n, d = a[0:4]
        return 4 * b * (2 * s * d + 2 * n * d) + 4 * b * s, 4 * b * s * n * d, "tensor"
    if name == "pasnl_nl_core_bwd":
        b, s, n, d = a[0:4]
        return 4 * b * (4 * s * d + 4 * n * d) + 8 * b * s, 10 * b * s * n * d, "tensor"
    if name == "pasnl_gemm_rows_tf32x3":
        rows, k, n = a[0:3]
        return 4 * rows * (k + n) + 4 * k * n, 2 * rows * k * n, gemm_bound(4 * rows * (k + n) + 4 * k * n, 2 * rows * k * n)
    if name in ("pasnl_gemm_wgrad_tf32x3", "pasnl_gemm_wgrad_tma"):
        rows, m, n = a[0:3]
        return 4 * rows * (m + n) + 4 * m * n, 2 * rows * m * n, gemm_bound(4 * rows * (m + n) + 4 * m * n, 2 * rows * m * n)
    if name == "pasnl_bn_act_fwd":       # apply reads y and writes out; + one more read of y when the statistics
        rows, c = a[0:2]                 # were not already produced by the GEMM epilogue (stats_ready = argument 15)
        stats_ready = bool(a[15]) if len(a) > 15 and isinstance(a[15], int) else False
        return (8 if stats_ready else 12) * rows * c, 8 * rows * c, "hbm"
    if name == "pasnl_bn_finalize":      # 2c doubles in, four c-vectors + the moving averages out
        c = a[1]
        return 16 * c + 4 * 8 * c, 12 * c, "hbm"
    if name == "pasnl_bn_act_bwd":       # reduce reads y, dout; apply reads y, dout, writes dy
        rows, c = a[0:2]
        return 20 * rows * c, 16 * rows * c, "hbm"
    if name in ("pasnl_gemm_rows_ws", "pasnl_gemm_rows_tma"):
        rows, k, n = a[0:3]
        return 4 * rows * (k + n) + 4 * k * n, 2 * rows * k * n, gemm_bound(4 * rows * (k + n) + 4 * k * n, 2 * rows * k * n)
    if name in ("pasnl_pointconv_fwd", "pasnl_pointconv_fwd_bn"):    # x (G,K,C) + w (G,K,32) in, (G,C,32) out
        g, k, c = a[0:3]
        return 4 * g * (k * c + 32 * k + 32 * c), 2 * g * k * c * 32, "hbm"
    if name in ("pasnl_pointconv_bwd", "pasnl_pointconv_bwd_bn"):    # x, w, grad_out in; grad_x, grad_w out
        g, k, c = a[0:3]
        return 4 * g * (2 * k * c + 2 * 32 * k + 32 * c), 4 * g * k * c * 32, "hbm"
    if name in ("pasnl_maxpool_k", "pasnl_maxpool_k_mask"):
        g, k, c = a[0:3]
        return 4 * g * k * c + (4 if name == "pasnl_maxpool_k" else 12) * g * c, g * k * c, "hbm"
    if name == "pasnl_maxpool_k_grad":   # x, out, grad_out in; dense grad_x out
        g, k, c = a[0:3]
        return 8 * g * k * c + 8 * g * c, g * k * c, "hbm"
    if name == "pasnl_maxpool_k_scatter":   # mask + grad_out in; ~one 32-byte sector read-modify-written per (g, c)
        g, k, c = a[0:3]
        return 12 * g * c + 64 * g * c, g * c, "hbm"
    if name == "pasnl_take_cols":        # reads rows x nc (strided), writes rows x nc
        rows, nc = a[0:2]
        return 8 * rows * nc, 0, "hbm"
    if name == "pasnl_transpose_batched":   # the matrix sizes live in a device table: charged one 32x32 tile per matrix,
        return 8 * 1024 * a[0], 0, "hbm"    # a LOWER bound (the true total is 8 bytes per weight element, 27 MB on cls)
    if name == "pasnl_add_cols":
        rows, nc = a[0:2]
        return 12 * rows * nc, rows * nc, "hbm"
    if name == "pasnl_add_n":
        n = a[0]
        return 4 * n * (4 if a[3] else 3), n * (2 if a[3] else 1), "hbm"
    if name in ("pasnl_concat_cols", "pasnl_split_cols"):
        rows, ca, cb = a[0:3]
        return 8 * rows * (ca + cb), 0, "hbm"
    if name in ("pasnl_bias_add", "pasnl_col_sum"):
        rows, n = a[0:2]
        return (8 if name == "pasnl_bias_add" else 4) * rows * n, rows * n, "hbm"
    if name == "pasnl_dropout":
        return 8 * a[0], 4 * a[0], "hbm"
    if name in ("pasnl_gemm_skinny",):
        rows, ni, nj = a[0:3]
        return 4 * (rows * ni + ni * nj + rows * nj), 2 * rows * ni * nj, "hbm"
    if name == "pasnl_wgrad_skinny":
        rows, ni, nj = a[0:3]
        return 4 * (rows * ni + rows * nj + ni * nj), 2 * rows * ni * nj, "hbm"
    if name in ("pasnl_softmax_ce_fwd", "pasnl_softmax_ce_bwd"):
        rows, c = a[0:2]
        return (4 if name.endswith("fwd") else 8) * rows * c + 16 * rows, 8 * rows * c, "hbm"
    if name == "pasnl_adam_l2":          # p, g, m, v read; p, m, v written; 1-byte mask
        return 29 * a[0], 12 * a[0], "hbm"
    return 0, 0, "hbm"


FP32_SIMT_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 148 SMs x 128 FMA lanes x 2 flop at the 1965 MHz the box runs


def roof_seconds(nbytes, flops, bound, peaks):
    """Time one launch would take AT its roof: max(bytes / HBM peak, flops / pipe peak).  3xTF32 runs three tensor
    passes at half the bf16 rate (peak/6); distance kernels are held to the fp32 SIMT peak; latency-bound ops (FPS) are
    only charged their compulsory bytes."""
    t = nbytes / (peaks["hbm_gbs"] * 1e9)
    if bound == "tensor":
        t = max(t, 6.0 * flops / (peaks["bf16_tflops"] * 1e12))
    elif bound == "fp32":
        t = max(t, flops / (FP32_SIMT_TFLOPS * 1e12))
    return t


class KernelProfiler:
    """CUDA events around every pasnl_* call (on torch's current stream, which is the stream the library launches
    on).  Used in a separate pass so that the headline timing carries no instrumentation."""

    def __init__(self):
        self.records = []
        self._start = None

    def before(self, name, args):
        self._start = torch.cuda.Event(enable_timing=True)
        self._start.record()

    def after(self, name, args):
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        ints = tuple(x if isinstance(x, (int, float)) else (None if x is None else 1) for x in args)
        self.records.append((name, ints, self._start, end))

    def table(self, steps):
        torch.cuda.synchronize()
        agg = {}
        for name, ints, s, e in self.records:
            ms = s.elapsed_time(e)
            by, fl, bound = algorithmic_work(name, ints)
            dims = []
            for v in ints:      # leading shape arguments; stop at the first device pointer
                if v is None or (isinstance(v, int) and abs(v) >= 2 ** 32):
                    break
                dims.append(round(v, 6) if isinstance(v, float) else int(v))
            key = (name,) + tuple(dims)
            r = agg.setdefault(key, {"name": name, "dims": list(key[1:]), "calls": 0, "ms": 0.0, "bytes": by,
                                     "flops": fl, "bound": bound, "_all": []})
            r["calls"] += 1
            r["_all"].append(ms)
        rows = []
        for r in agg.values():
            # MEDIAN launch duration x calls: the eager pass is CPU-bound, so an event pair occasionally also spans a host
            # hiccup (allocator, GC) between the start event and the launch -- one such outlier must not pick the kernel
            samples = sorted(r.pop("_all"))
            avg = samples[len(samples) // 2]
            r["ms"] = avg * r["calls"]
            rows.append({**r, "avg_us": 1e3 * avg, "ms_per_step": r["ms"] / steps,
                         "roof_ms_per_step": 1e3 * roof_seconds(r["bytes"], r["flops"], r["bound"], load_peaks()) * r["calls"] / steps,
                         "gbs": r["bytes"] / (avg * 1e-3) / 1e9 if avg > 0 else 0.0,
                         "tflops": r["flops"] / (avg * 1e-3) / 1e12 if avg > 0 else 0.0})
        rows.sort(key=lambda r: -r["ms_per_step"])
        return rows


# ------------------------------------------------------------------------------------------------ reference arm
def run_cpu_reference(args, batch, points, steps, warmup, budget_s=40.0, full_batch=False):
    """The reference's algorithm on host cores: oracle/model_ref.py (line-by-line restatement of the TF graph code,
    reference-native nanoflann kNN / three_nn / three_interpolate where oracle/_ref has them, C restatement for the
    GPU-only ops).  Bounded sample: `sample_batch` clouds per step (per-cloud work is independent apart from the
    batch-norm statistics, so clouds/s scales linearly), thread count chosen by a short scan because the 1x1-conv
    stack on 10^5-row tensors scales badly past a few dozen threads.
    Returns (clouds_per_s, ms_per_step, threads_used, kind, sample description, host core count)."""
    from oracle import cref, model_ref, refnative
    cores = os.cpu_count() or 1
    # the in-line cpu_baseline leg times a bounded sample; `--impl reference` (full_batch) runs the whole cls batch
    sample_batch = min(batch, 8 if args.workload == "cls" else (2 if args.workload == "seg" else 1))
    if full_batch:
        sample_batch = batch
    data = {}

    def load(nb):
        pc, labels, smpw = make_batch(args.workload, nb, points, 1)
        data["x"], data["y"] = torch.from_numpy(pc), torch.from_numpy(labels)
        data["w"] = torch.from_numpy(smpw) if smpw is not None else None

    load(sample_batch)
    r = model_ref.Ref(None, dtype=torch.float32, dropout=True, native_knn=True, native_interp=True)

    def fwd():
        if args.workload == "cls":
            return model_ref.cls_forward(r, data["x"], data["y"], use_normal=True, adaptive_sample=args.AS)[1]
        kw = WORKLOADS[args.workload]["kw"]
        return model_ref.seg_forward(r, data["x"], kw["num_class"], data["y"], data["w"],
                                     feature_channel=kw["feature_channel"], radius=kw["radius"])[1]

    torch.set_num_threads(min(cores, 16))
    with torch.no_grad():
        fwd()          # instantiate the variables (lazily, with the reference's initialisers)
    params = [v for v in r.v.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)

    def one_step():
        opt.zero_grad(set_to_none=True)
        total = fwd()
        total.backward()
        opt.step()
        return float(total.detach())

    # thread scan: one step each, keep the fastest setting ("all the host threads it can use" -- more is slower)
    t_start = time.perf_counter()
    best = None
    for th in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(th)
        cref.set_threads(min(th, sample_batch))
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (th, dt)
        if time.perf_counter() - t_start > budget_s / 2:
            break
    threads = best[0]
    torch.set_num_threads(threads)
    # `--impl reference --steps K --warmup W` runs exactly K + W steps: when that would not fit the time budget at the
    # full batch, the per-step SAMPLE shrinks (clouds are independent apart from the batch-norm statistics), not K
    if full_batch and (steps + warmup) * best[1] > budget_s and sample_batch > 2:
        sample_batch = max(2, int(sample_batch * budget_s / ((steps + warmup) * best[1])))
        load(sample_batch)
    cref.set_threads(min(threads, sample_batch))
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(1, steps)):
        one_step()
        done += 1
        if not full_batch and time.perf_counter() - t0 > budget_s / 2:
            break
    dt = time.perf_counter() - t0
    kind = "reference" if refnative.available("ref_knn.so") else "port"
    sample = (f"{done} step(s) of {sample_batch} clouds (of the {batch}-cloud batch) after a thread scan "
              f"(best of 8/16/32/64/{cores}: {threads} threads); torch CPU + reference-native nanoflann kNN / three_nn / "
              "three_interpolate (oracle/_ref), FPS and gathers from the C restatement (the reference has no CPU FPS)")
    return sample_batch * done / dt, 1e3 * dt / done, threads, kind, sample, cores


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    batch = args.batch or wl["batch"]
    points = args.points or wl["points"]
    global_batch = batch * world
    if args.scaling == "strong":          # --batch clouds in total; rank r takes dp.shard_range(batch, r, world)
        from pointasnl_b200 import dp as _dp
        global_batch = batch
        lo, hi = _dp.shard_range(global_batch, rank, world)
        batch = hi - lo
        if batch < 2:
            raise SystemExit(f"strong scaling: rank {rank} would get {batch} cloud(s); batch norm needs >= 2 per replica")
    model_name = {"cls": "pointasnl_cls" + ("+AS" if args.AS else ""), "seg": "pointasnl_sem_seg_res (ScanNet grid, rgb)",
                  "kitti": "pointasnl_sem_seg_res (SemanticKITTI grid, remission)"}[args.workload]
    workload_name = model_name + (f" B={batch} N={points} per GPU" if args.scaling == "weak" else
                                  f" global B={global_batch} N={points} sharded over {world} GPU(s)") + ", synthetic"
    metric = METRIC if args.workload == "cls" else f"point-clouds/sec fwd+bwd pointasnl_sem_seg_res B={wl['batch']} N={points}"
    # one protocol description shared by both arms (so that the two JSON lines carry the same `config`)
    config = {"workload": workload_name, "step": "forward + loss + backward + grad all-reduce + Adam",
              "parallelism": f"dp{world}", "global_batch": global_batch,
              "batch_norm": "per-replica statistics" if world > 1 else "single replica",
              "l2": "GPU arm: flushed (256 MiB memset) before every timed step, outside the event pair",
              "launch": "GPU arm: CUDA graphs (fwd+bwd, optimizer) around one NCCL all-reduce, or eager with --no-graph; "
                        "reference arm: host threads"}

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        # exactly --steps / --warmup steps of the same workload on the host cores; the per-step sample is the full batch
        # when that fits ~3 minutes, a bounded share of it otherwise
        steps, warm = max(1, args.steps), max(0, args.warmup)
        val, ms, threads, kind, sample, cores = run_cpu_reference(args, batch, points, steps, warm, budget_s=180.0,
                                                                  full_batch=True)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "threads": threads, "kind": kind,
                                 "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- this repo's CUDA path
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from pointasnl_b200 import _lib, dp
    from pointasnl_b200.callers.train_step import TrainStep

    pc, labels, smpw = make_batch(args.workload, batch, points, 1 if world == 1 else 1000 + rank)
    h_pc = torch.from_numpy(pc).pin_memory()
    h_lab = torch.from_numpy(labels).pin_memory()
    h_w = torch.from_numpy(smpw).pin_memory() if smpw is not None else None
    ts = TrainStep(wl["model"], batch, points, dev, adaptive_sample=args.AS, use_graph=not args.no_graph, seed=0, **wl["kw"])
    ts.set_inputs(h_pc, h_lab, h_w)
    torch.cuda.synchronize()
    if not args.no_graph:
        ts.capture(warmup=3)
    for _ in range(max(args.warmup, 3)):
        ts.step()
    torch.cuda.synchronize()
    launches_per_step = None
    if args.no_graph:
        c0 = _lib.launch_count
        ts.step()
        launches_per_step = _lib.launch_count - c0
    else:
        launches_per_step = ts.launches_per_step

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for i in range(args.steps):
            flush.zero_()
            starts[i].record()
            ts.step()
            ends[i].record()
        barrier()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = dp.max_over_ranks(sum(step_ms), dev)
    value = global_batch * args.steps / (total_ms * 1e-3)

    # ---- timed region 2: end to end from pinned host memory, loss read back every step
    e2e_t = []
    barrier()
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts.set_inputs(h_pc, h_lab, h_w)
        loss_host = float(ts.step().item())     # D2H read of the step's result
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    assert np.isfinite(loss_host)
    e2e_total = dp.max_over_ranks(sum(e2e_t), dev)
    e2e_value = global_batch * args.steps / e2e_total
    h2d = h_pc.numel() * 4 + h_lab.numel() * 8 + (h_w.numel() * 4 if h_w is not None else 0)

    if world > 1:
        # every rank leaves the process group TOGETHER, and before rank 0's single-process passes: the captured graphs
        # (which hold NCCL kernels) are released first
        torch.cuda.synchronize()
        torch.distributed.barrier()
        launch_mode_single_graph = getattr(ts, "g_all", None) is not None
        ts.g_all = ts.g_fb = ts.g_opt = None
        ts.graph = None
        torch.cuda.synchronize()
        torch.distributed.destroy_process_group()
    else:
        launch_mode_single_graph = False
    if rank != 0:
        return

    # ---- per-kernel pass (eager, CUDA events around each pasnl launch) -> roofline of the dominant kernel
    peaks = load_peaks()
    prof_steps = min(args.steps, 5)
    # the per-launch pass runs with the side / branch streams OFF: every launch is then alone on the device between its
    # two events (with them on, a launch's events also span the time it shares SMs with the other branch)
    from pointasnl_b200 import pointasnl_util
    pointasnl_util.set_concurrency(False)
    ts_prof = None
    if ts_prof is None:
        ts_prof = TrainStep(wl["model"], batch, points, dev, adaptive_sample=args.AS, use_graph=False, seed=0,
                            overlap_allreduce=False, **wl["kw"])
        ts_prof.bucket.all_reduce_sum = lambda *a, **k: None    # rank-0 only pass: no collective
        ts_prof.set_inputs(h_pc, h_lab, h_w)
        for _ in range(2):
            ts_prof.step()
    prof = KernelProfiler()
    torch.cuda.synchronize()
    _lib.profiler = prof
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(prof_steps):
        ts_prof.step()
    ev1.record()
    _lib.profiler = None
    pointasnl_util.set_concurrency(True)
    rows = prof.table(prof_steps)
    eager_ms = ev0.elapsed_time(ev1) / prof_steps
    ours_ms = sum(r["ms_per_step"] for r in rows)
    # dominant kernel = the entry point with the largest share of the step, reported on the shape it spends most time on
    # (the rows GEMM is one op with two C entry points: pasnl_gemm_rows_tf32x3 routes tall layers to pasnl_gemm_rows_ws and
    #  the Python layer calls the latter directly when it wants the [k][n] weight layout -- they are ranked together)
    def op_of(name):
        if name in ("pasnl_gemm_rows_tf32x3", "pasnl_gemm_rows_ws", "pasnl_gemm_rows_tma"):
            return "pasnl_gemm_rows"
        return "pasnl_gemm_wgrad" if name in ("pasnl_gemm_wgrad_tf32x3", "pasnl_gemm_wgrad_tma") else name
    # Farthest point sampling is a latency chain (npoint dependent rounds on one CTA / cluster per cloud, a few KB of
    # traffic): it has no bandwidth or pipe roofline and, at SemanticKITTI size, runs hidden on a side stream beside the
    # previous layer.  It is reported next to the roofline (`latency_bound_ops`), not ranked as the dominant kernel.
    by_name, latency_ops = {}, {}
    for r in rows:
        if r["bound"] == "latency":
            e = latency_ops.setdefault(r["name"], {"ms_per_step": 0.0, "largest": None})
            e["ms_per_step"] += r["ms_per_step"]
            if e["largest"] is None or r["avg_us"] > e["largest"]["avg_us"]:
                e["largest"] = {"dims": r["dims"], "avg_us": r["avg_us"]}
                if r["name"] == "pasnl_fps" and r["avg_us"] > 0 and len(r["dims"]) >= 3:
                    b_, n_, m_ = (float(v) for v in r["dims"][0:3])  # rounds/s and the on-chip sweep rate
                    e["largest"]["rounds_per_s"] = m_ / (r["avg_us"] * 1e-6)
                    e["largest"]["onchip_sweep_gbs"] = 16.0 * b_ * m_ * n_ / (r["avg_us"] * 1e-6) / 1e9
            continue
        by_name[op_of(r["name"])] = by_name.get(op_of(r["name"]), 0.0) + r["ms_per_step"]
    top_name = max(by_name, key=by_name.get)
    top = max((r for r in rows if op_of(r["name"]) == top_name), key=lambda r: r["ms_per_step"])
    if top["bound"] == "tensor":
        roof = {"bound": "tensor", "achieved": top["tflops"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": top["tflops"] / peaks["bf16_tflops"]}
    elif top["bound"] == "fp32":
        # the brute-force distance kernels (k-NN / ball query / three_nn at SemanticKITTI size) are bound by the fp32 FMA
        # pipe, not by HBM or the tensor cores: reported against the SIMT peak (148 SMs x 128 lanes x 2 x 1965 MHz)
        roof = {"bound": "fp32", "achieved": top["tflops"], "peak": FP32_SIMT_TFLOPS, "unit": "TFLOP/s",
                "frac": top["tflops"] / FP32_SIMT_TFLOPS}
    else:
        roof = {"bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": top["gbs"] / peaks["hbm_gbs"]}
    traffic = load_ncu_traffic().get(top["name"] + "|" + ",".join(str(d) for d in top["dims"][:3]))
    roof.update({"traffic": traffic, "kernel": top["name"], "dims": top["dims"],
                 "avg_us": top["avg_us"], "entry_point_ms_per_step": by_name[top_name],
                 "native_bound": top["bound"], "peak_source": peaks["src"], "latency_bound_ops": latency_ops,
                 "timing": f"CUDA events around each launch (median per shape), eager pass of {prof_steps} steps, side/branch streams off",
                 # sum of the event-timed library launches vs the wall time of the same EAGER step (the eager step is
                 # CPU-launch bound, so the ratio understates the share; under ncu the library is 88 % of the device
                 # time of a step: profiles/r01c_launches_cls_step.csv)
                 "pasnl_ms_per_step_eager_events": ours_ms, "eager_ms_per_step": eager_ms})
    # whole-step hot-path roofline: every library launch at ITS roof, against the graph-replayed step
    ideal_ms = sum(r["roof_ms_per_step"] for r in rows)
    by_roof = {}
    for r in rows:
        e = by_roof.setdefault(r["name"], [0.0, 0.0])
        e[0] += r["ms_per_step"]
        e[1] += r["roof_ms_per_step"]
    step_roof = {"ideal_ms_per_step": ideal_ms, "clouds_per_s_at_roof": batch / (ideal_ms * 1e-3) if ideal_ms > 0 else None,
                 "measured_ms_per_step": total_ms / args.steps, "frac": ideal_ms / (total_ms / args.steps),
                 "definition": "sum over library launches of max(algorithmic bytes / HBM peak, flops / pipe peak); "
                               "3xTF32 tensor peak = bf16 peak / 6, fp32 SIMT peak = %.1f TFLOP/s" % FP32_SIMT_TFLOPS,
                 "per_op": [{"op": k, "ms": round(v[0], 4), "roof_ms": round(v[1], 4),
                             "frac": round(v[1] / v[0], 3) if v[0] > 0 else None}
                            for k, v in sorted(by_roof.items(), key=lambda kv: -kv[1][0])[:14]]}
    if args.kernel_report:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernel_report)), exist_ok=True)
        json.dump({"eager_ms_per_step": eager_ms, "pasnl_ms_per_step": ours_ms, "peaks": peaks, "kernels": rows},
                  open(args.kernel_report, "w"), indent=1)

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, ms, threads, kind, sample, cores = run_cpu_reference(args, batch, points, steps=2, warmup=0, budget_s=30.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "threads": threads, "kind": kind, "sample": sample}
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}

    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config,
            "launch_mode": "eager" if args.no_graph else (
                "one CUDA graph per step with a two-piece overlapped all-reduce" if launch_mode_single_graph else
                "CUDA graphs (fwd+bwd, optimizer) around one NCCL all-reduce"),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "step_roofline": step_roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
