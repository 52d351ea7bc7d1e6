#!/usr/bin/env python
"""tools/ncu_summary.py — turn ncu output into the small text summaries kept under profiles/.

    python tools/ncu_summary.py rep   gpurun_out/x.ncu-rep   > profiles/x.json      # --set full capture
    python tools/ncu_summary.py list  gpurun_out/launches.csv > profiles/x.json     # launch list

`rep` keeps the handful of raw metrics the roofline argument needs (duration, DRAM bytes,
throughput percentages, shared-memory wavefronts / bank conflicts, issue utilisation, stall
reasons, tensor-pipe activity, registers) per profiled launch.
`list` aggregates the per-launch gpu__time_duration list by kernel name: count, total, mean and
share of the summed device time (per-launch times under ncu are cold-cache and serialised, so
the SHARE is what is comparable with the CUDA-event numbers of bench.py).
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "Kernel Name", "Block Size", "Grid Size",
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def rep(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {}
        for k in KEEP:
            if k in idx and r[idx[k]] != "":
                d[k] = (r[idx[k]] + " " + units[idx[k]]).strip()
        out.append(d)
    return {"source": path, "launches": out}


def launch_list(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    unit = ""
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[idx["Metric Unit"]]
        a = agg[r[idx["Kernel Name"]].split("(")[0][-60:]]
        a[0] += 1
        a[1] += float(r[idx["Metric Value"]].replace(",", ""))
    tot = sum(v[1] for v in agg.values()) or 1.0
    ours = ("gather_", "dense_tc", "code_gemv", "split_tiles", "tile_codes", "pad_lists", "scan_counts", "rebase",
            "emit_indices", "encode_planes", "padded_counts", "planes_from", "scatter_dense",
            "pcsc", "tcsr", "codes_")
    ks = [{"kernel": k, "launches": n, "total": round(t, 1), "mean": round(t / n, 2),
           "share": round(t / tot, 4), "ours": any(o in k for o in ours)}
          for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    return {"source": path, "unit": unit, "total": round(tot, 1), "kernels": ks}


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    json.dump(rep(path) if mode == "rep" else launch_list(path), sys.stdout, indent=1)
    print()
