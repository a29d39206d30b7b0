#!/usr/bin/env python3
"""Writes profiles/count_kernel_traffic.json from `ncu --set full` captures: per capture the DRAM bytes of ONE launch
(dram__bytes_read.sum + dram__bytes_write.sum), the kernel it is of, and a few memory-system metrics — stamped with
a hash of the kernel sources so that bench.py only uses a capture that belongs to the code it is running.

usage: tools/ncu_traffic.py key=report.ncu-rep[:kernel-substring[:units]] ...      (keys: c3, c3_stepping, c3_large_table, c2, c5, c4_walk)
`units` (optional) = work items of the captured launch (occurrences of the walk): bench.py scales the bytes to its own batch.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KEEP = ["gpu__time_duration.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def sources_sha16():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".hpp")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def read_report(path, want):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        if want and want not in name:
            continue

        def val(k):
            i = hdr.index(k)
            return float(row[i].replace(",", "")), units[i]

        rd, u1 = val("dram__bytes_read.sum")
        wr, u2 = val("dram__bytes_write.sum")
        ent = {"kernel": name, "dram_bytes_per_launch": rd * UNITS[u1] + wr * UNITS[u2], "report": os.path.basename(path)}
        for k in KEEP:
            if k in hdr:
                v, u = val(k)
                ent[k] = v if not u else f"{v:g} {u}"
        req = ent.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
        sec = ent.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        try:
            ent["sectors_per_request"] = float(str(sec).split()[0]) / float(str(req).split()[0])
        except Exception:
            pass
        return ent
    raise SystemExit(f"{path}: no kernel matching {want!r}")


def main():
    out = {"source_sha16": sources_sha16(), "generator": "tools/ncu_traffic.py", "captures": {}}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        parts = rest.split(":")
        path, want = parts[0], parts[1] if len(parts) > 1 else ""
        out["captures"][key] = read_report(path, want)
        if len(parts) > 2:
            out["captures"][key]["units_per_launch"] = int(parts[2])
    json.dump(out, open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
