"""Summarise one kernel of an .ncu-rep: headline metrics, pipe utilisation, stall reasons, and
(with --source) the hottest SASS instructions.  Reads the report with `ncu -i`; no GPU needed."""
import csv, subprocess, sys, io, argparse
from collections import Counter
ap = argparse.ArgumentParser(); ap.add_argument("rep"); ap.add_argument("--source", action="store_true"); ap.add_argument("--top", type=int, default=25)
ap.add_argument("--range", default=None, help="a:b SASS line range to print")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    st = [(float(d[h]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and d[h]]
    st.sort(reverse=True)
    print("  stall cycles per issued instruction: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:10]))
if a.source:
    src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]; data = rows[2:]
    iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[iE]) for r in data); tw = sum(int(r[iW]) for r in data)
    print(f"  SASS: {len(data)} instructions, {tot} executed (warp level), {tw} stall samples")
    c = Counter(); cs = Counter()
    for r in data:
        parts = r[iS].split(); op = parts[1] if parts[0].startswith("@") else parts[0]
        c[op] += int(r[iE]); cs[op] += int(r[iW])
    for op, n in c.most_common(a.top):
        print(f"    {op:30s} {n:11d} {100*n/tot:6.2f}%   samples {100*cs[op]/max(tw,1):6.2f}%")
    if a.range:
        lo, hi = map(int, a.range.split(":"))
        for i, r in enumerate(data[lo:hi]):
            print(f"{i+lo:5d} {r[iS].strip()[:84]:84s} {r[iE]:>9s} {r[iW]:>5s}")
