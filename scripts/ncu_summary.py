"""Summarise an `ncu --set full --import-source on` capture of one kernel for profiles/.

usage: ncu_summary.py REPORT.ncu-rep > profiles/rNN_ncu_<kernel>.txt
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source cuda,sass`.
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]


def page(args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + args, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page(["--page", "raw"])
hdr, units, val = raw[0], raw[1], raw[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
print(f"# {rep}")
print(f"kernel: {d.get('Kernel Name', ('?',))[0]}   grid {d.get('Grid Size', ('?',))[0]} block {d.get('Block Size', ('?',))[0]}")
keys = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:>16s} {d[k][1]}")
print("\n# warp stalls per issued instruction")
st = {k: float(v[0]) for k, v in d.items()
      if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio")}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:6.3f}")

src = page(["--page", "source", "--print-source", "cuda,sass"])
cur = None
line = None
per_line = {}
ops = Counter()
for r in src:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] == "Line No":
        continue
    if r[0] != "":
        try:
            line = (cur, int(r[0]), r[1])
        except ValueError:
            pass
        continue
    try:
        n, smp = int(r[7]), int(r[4])
    except ValueError:
        continue
    a = per_line.setdefault(line[:2], [line[2], 0, 0])
    a[1] += n
    a[2] += smp
sass = page(["--page", "source", "--print-source", "sass"])
sh = sass[1]
ci, cs = sh.index("Instructions Executed"), sh.index("Source")
for r in sass[2:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    tok = r[cs].split()
    if tok:
        op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
        ops[op.split(".")[0]] += int(r[ci])
tot = sum(a[1] for a in per_line.values()) or 1
otot = sum(ops.values()) or 1
tots = sum(a[2] for a in per_line.values()) or 1
print(f"\n# executed warp instructions by opcode (total {otot})")
for op, n in ops.most_common(24):
    print(f"  {op:10s} {100 * n / otot:5.2f} %")
print("\n# hottest source lines (share of executed instructions, share of stall samples)")
for (f, l), a in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"  {f}:{l:<4d} {100 * a[1] / tot:5.2f} %  {100 * a[2] / tots:5.2f} %  {a[0].strip()[:100]}")
