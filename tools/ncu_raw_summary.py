"""Print the key columns of an `ncu --page raw --csv` export, one line per profiled launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__inst_executed.sum", "inst")]
for r in rows[2:]:
    out = []
    for c, n in cols:
        if c in idx:
            v = r[idx[c]]
            if n == "kernel":
                v = v.split("(")[0][-60:]
            else:
                try:
                    v = "%.3g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                v += units[idx[c]] if n in ("rd", "wr", "ms") else ""
            out.append("%s=%s" % (n, v))
    print(" ".join(out))
