#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> [out.csv] — per-launch summary of an `ncu --set full` report for profiles/:
duration, DRAM bytes and % of peak, issue-slot %, achieved occupancy, registers, warp instructions, shared/global atomic
traffic and the five largest warp-stall reasons (share of sampled warp states). Run on the CPU box (ncu -i needs no GPU)."""
import csv
import io
import subprocess
import sys

COLS = [("kernel", "Kernel Name"), ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
        ("dur_us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("warp_inst", "smsp__inst_executed.sum"),
        ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"), ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smem_atom_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"),
        ("smem_ld_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
        ("global_atom_sectors", "lts__t_sectors_op_atom.sum"), ("global_red_sectors", "lts__t_sectors_op_red.sum")]


def to_unit(v, unit, want):
    try:
        x = float(v.replace(",", ""))
    except Exception:
        return v
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
    bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    if want == "us" and unit in scale:
        return round(x * scale[unit], 3)
    if want == "MB" and unit in bscale:
        return round(x * bscale[unit], 3)
    return round(x, 3)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    out = []
    for r in body:
        d = {}
        for name, col in COLS:
            if col not in hdr:
                d[name] = ""
                continue
            i = hdr.index(col)
            want = "us" if name == "dur_us" else ("MB" if name.endswith("_MB") else "")
            d[name] = to_unit(r[i], units[i], want) if name != "kernel" else r[i].split("(")[0].replace("void ", "")[:60]
        tot = 0.0
        vals = []
        for i, h in stall:
            try:
                v = float(r[i].replace(",", ""))
            except Exception:
                v = 0.0
            vals.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            tot += v
        vals.sort(reverse=True)
        d["top_stalls"] = " ".join(f"{n}:{100 * v / tot:.0f}%" for v, n in vals[:5]) if tot else ""
        out.append(d)
    f = open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout
    w = csv.DictWriter(f, fieldnames=[c[0] for c in COLS] + ["top_stalls"])
    w.writeheader()
    w.writerows(out)


if __name__ == "__main__":
    main()
