"""Runs a few warm-up training steps, then exactly `--steps` steps inside a cudaProfilerStart/Stop window, so that
ncu (--profile-from-start off) and torch.profiler see only steady-state steps of the benchmark workload.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/one_step.py
  python tools/one_step.py --torch-profiler gpurun_out/torch_prof.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS, make_batch  # noqa: E402
from pointasnl_b200.callers.train_step import TrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cls")
    ap.add_argument("--AS", action="store_true")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--torch-profiler", default=None)
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    batch = a.batch or wl["batch"]
    points = a.points or wl["points"]
    dev = torch.device("cuda:0")
    pc, labels, smpw = make_batch(a.workload, batch, points, 1)
    ts = TrainStep(wl["model"], batch, points, dev, adaptive_sample=a.AS, use_graph=False, **wl["kw"])
    ts.set_inputs(torch.from_numpy(pc), torch.from_numpy(labels), None if smpw is None else torch.from_numpy(smpw))
    for _ in range(3):
        ts.step()
    torch.cuda.synchronize()
    if a.torch_profiler:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(a.steps):
                ts.step()
            torch.cuda.synchronize()
        with open(a.torch_profiler, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
        return
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        ts.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
