#!/usr/bin/env python
"""Condense an `.ncu-rep` (read here with `ncu -i ... --page raw --csv`, no GPU needed) into the per-kernel JSON summaries
committed under profiles/:  python tools/ncu_summary.py gpurun_out/r02_tc.ncu-rep profiles/r02_ncu_full_tc_summary.json "<command>"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "nvlrx__bytes.sum", "nvltx__bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        name = d.get("Kernel Name", "?")
        m = {}
        for k, u in zip(head, units):
            if k in KEEP and d.get(k, "") != "":
                m[k] = {"value": d[k], "unit": u}
        kernels.append({"kernel": name, "id": d.get("ID"), "metrics": m})
    with open(out, "w") as f:
        json.dump({"report": rep, "command": cmd, "kernels": kernels}, f, indent=1)
    for k in kernels:
        g = lambda n: k["metrics"].get(n, {}).get("value", "-")
        print(k["kernel"][:70], "| ms", g("gpu__time_duration.sum"), "| tensor%", g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
              "| dram R/W", g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), "| GHz", g("sm__cycles_elapsed.avg.per_second"))


if __name__ == "__main__":
    main()
