#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: a launch list (CSV from --metrics gpu__time_duration.sum) -> per-kernel totals,
and a --set full report (.ncu-rep) -> key metrics per captured kernel.
    python tools/ncu_summary.py launches <csv>
    python tools/ncu_summary.py report <ncu-rep>
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
]  # fmt: skip


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = row.get("Metric Unit", "")
        v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        name = re.sub(r"\(.*", "", row.get("Kernel Name", ""))
        if name.startswith("void at::") or name.startswith("void at_cuda") or "<unnamed>" in name:
            name = "torch: " + name[:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel totals of one step under ncu (serialised, cold cache): total {tot:.3f} ms")
    print(f"{'kernel':64s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:64s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot:7.3f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("##", r[ik][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:82s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
