#!/usr/bin/env python3
"""Key metrics of every kernel launch in an .ncu-rep (from `ncu --set full`) as JSON.

usage: tools/ncu_summary.py report.ncu-rep [kernel-name-substring] > profiles/<name>.json
Stall shares come from smsp__average_warps_issue_stalled_<reason>_per_issue_active.ratio.
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        if want and want not in name:
            continue
        d = {"Kernel Name": name}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{row[i]} {units[i]}".strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX) and "not_issued" not in h:
                try:
                    stalls[h[len(STALL_PREFIX):-len(STALL_SUFFIX)]] = float(row[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values())
        if tot > 0:
            d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot >= 0.01}
        out.append(d)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
