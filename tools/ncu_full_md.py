"""Markdown table of an `ncu --set full` capture (raw csv page), one row per distinct kernel.

    python tools/ncu_full_md.py <raw.csv> <trajectories per launch by kernel-substring, e.g. "step_kernel<double, 7, 6, 2, 0=32,default=16"> <points per trajectory>
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
spec = dict(kv.rsplit("=", 1) for kv in sys.argv[2].split(";"))
pts_traj = float(sys.argv[3])


def f(r, c):
    return float(r[idx[c]].replace(",", ""))


def scale(r, c):  # bytes in whatever unit ncu chose -> bytes
    u = rows[1][idx[c]]
    return f(r, c) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def ms(r):
    u = rows[1][idx["gpu__time_duration.sum"]]
    return f(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3,
                                              "msecond": 1}[u]


seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    seen.setdefault(name, []).append(r)
print("| kernel | trajectories / launch | ms | fp64 pipe % (active) | issue slots % | smem wavefronts % | regs | "
      "warps active % | DRAM read GB | DRAM write GB | DRAM B/point | warp instr |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, rs in seen.items():
    r = rs[len(rs) // 2]
    ntraj = float(next((v for k, v in spec.items() if k != "default" and k in name), spec["default"]))
    rd, wr = scale(r, "dram__bytes_read.sum"), scale(r, "dram__bytes_write.sum")
    print("| `%s` | %d | %.3f | %.1f | %.1f | %.1f | %d | %.1f | %.3f | %.3f | %.0f | %.3g |" % (
        name, ntraj, ms(r), f(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        f(r, "launch__registers_per_thread"), f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        rd / 1e9, wr / 1e9, (rd + wr) / (ntraj * pts_traj), f(r, "smsp__inst_executed.sum")))
