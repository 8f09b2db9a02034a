"""Condenses an .ncu-rep (ncu --set full) into a small markdown table under profiles/ (run on the CPU box).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_xxx.md "title"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    lines = ["# %s" % title, "", "Source: `%s` (ncu --set full --clock-control none, one launch each, cold-cache/serialised: compare shares, not absolutes)." % rep, ""]
    names = [r[idx["Kernel Name"]].split("(")[0][:40] for r in kernels]
    lines.append("| metric | " + " | ".join(names) + " |")
    lines.append("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in idx:
            continue
        vals = []
        for r in kernels:
            v = r[idx[key]]
            try:
                x = float(v.replace(",", ""))
                v = "%.4g" % x
            except ValueError:
                pass
            vals.append("%s %s" % (v, units[idx[key]]))
        lines.append("| %s | " % label + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu summary")
