"""Markdown summary of an ncu report: one row per captured launch with the metrics the profiles/ files quote.
    python tools/ncu_summary.py <report.ncu-rep> [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
M = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
     ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"), ("smsp__inst_executed.sum", "warp instr"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("lts__t_sectors_srcunit_ltcfabric.sum", "L2 fabric sectors"),
     ("nvlrx__bytes.sum", "NVLink rx"), ("nvltx__bytes.sum", "NVLink tx")]
idx = [(lab, hdr.index(m)) for m, lab in M if m in hdr]
print("| " + " | ".join(l + (" (" + units[i] + ")" if units[i] else "") for l, i in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    if sub and sub not in r[hdr.index("Kernel Name")]:
        continue
    cells = []
    for lab, i in idx:
        v = r[i]
        if lab == "kernel":
            v = v.split("(")[0].replace("void ", "")[:48]
        else:
            try:
                f = float(v.replace(",", ""))
                v = ("%.4g" % f) if abs(f) < 1e6 else ("%.4e" % f)
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
