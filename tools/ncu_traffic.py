"""Runs `ncu --set full` on single launches of the dominant kernels (tools/ncu_ops.py) and writes
gpurun_out/ncu_traffic.json: per entry point + shape, DRAM bytes read / written, duration, L1/shared data-pipe and
tensor-pipe utilisation.  Copy the result to profiles/ncu_traffic.json (bench.py reads `roofline.traffic` from it) and
the raw CSVs to profiles/.   python tools/ncu_traffic.py [op ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ncu_ops import OPS  # noqa: E402

METRICS = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
           "gpu__time_duration.sum": "duration_ns", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "launch__registers_per_thread": "registers",
           "sm__inst_executed_pipe_tensor.sum": "tensor_insts", "smsp__cycles_active.avg": "smsp_cycles_active",
           "l1tex__data_bank_conflicts_pipe_lsu.sum": "lsu_bank_conflicts",
           "smsp__inst_executed.sum": "inst_executed", "sm__cycles_elapsed.max": "sm_cycles"}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3,
              "msecond": 1e6}


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    names = sys.argv[1:] or list(OPS)
    result = {"traffic": {}, "kernels": {}}
    for name in names:
        entry, dims, regex, _ = OPS[name]
        rep = os.path.join(out_dir, f"ncu_{name}")
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{regex}",
               "--launch-skip", "2", "-c", "1", "-f", "-o", rep, sys.executable, os.path.join(ROOT, "tools", "ncu_ops.py"), name]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(rep + ".ncu-rep"):
            print(name, "ncu failed:", r.stderr[-500:])
            continue
        raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(out_dir, f"ncu_{name}_raw.csv"), "w").write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        rec = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else regex}
        for m, key in METRICS.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    rec[key] = float(vals[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1)
                except ValueError:
                    pass
        os.remove(rep + ".ncu-rep") if os.path.getsize(rep + ".ncu-rep") > 30 * 2 ** 20 else None
        result["kernels"][name] = {"entry_point": entry, "dims": dims, **rec}
        if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
            result["traffic"][entry + "|" + ",".join(str(d) for d in dims[:3])] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
        print(name, json.dumps(rec))
    json.dump(result, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
