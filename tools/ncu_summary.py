#!/usr/bin/env python3
"""Summarise an ncu report (read here, no GPU needed): key raw metrics + top stall sites.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_csr_merge_warp.txt [traffic_key]
When traffic_key is given, profiles/traffic.json[traffic_key] = dram read + write bytes per launch."""
import csv
import io
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]

UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    lines = ["ncu summary of %s" % os.path.basename(rep), ""]
    traffic = None
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append("kernel: %s" % d.get("Kernel Name", "?"))
        for w in WANT:
            if w in d:
                lines.append("  %-66s %18s %s" % (w, d[w], u[w]))
        try:
            rd = float(d["dram__bytes_read.sum"]) * UNIT_SCALE[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * UNIT_SCALE[u["dram__bytes_write.sum"]]
            traffic = rd + wr
            lines.append("  %-66s %18.0f byte" % ("dram traffic (read + write) per launch", traffic))
        except Exception:
            pass
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    if len(src) > 2:
        h = src[1]
        idx = {n: i for i, n in enumerate(h)}
        data = [r for r in src[2:] if len(r) == len(h)]
        if "# Samples" in idx:
            stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
            tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
            agg = sorted(((sum(int(r[idx[s]] or 0) for r in data), s) for s in stalls), reverse=True)
            lines.append("warp stall samples (all): total %d" % tot)
            for n, s in agg[:8]:
                lines.append("  %-28s %8d  %5.1f%%" % (s, n, 100.0 * n / max(tot, 1)))
            lines.append("")
            lines.append("top SASS sites by samples:")
            for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:16]:
                st = {s: int(r[idx[s]] or 0) for s in stalls}
                lines.append("  %8s  %-22s %s" % (r[idx["# Samples"]], max(st, key=st.get), r[idx["Source"]][:90]))
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))
    if key and traffic:
        tpath = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
        t = json.load(open(tpath)) if os.path.exists(tpath) else {}
        t[key] = traffic
        json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
