"""Summarise an .ncu-rep (raw page + per-opcode instruction mix from the source page)."""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:90])
    for k in keys:
        if k in hdr:
            print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued")]
    tot = sum(float(v) for _, v in st if v.replace('.', '').isdigit())
    for h, v in sorted(st, key=lambda t: -float(t[1]) if t[1].replace('.', '').isdigit() else 0)[:9]:
        print("  stall %-40s %5.1f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * float(v) / tot))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, s = Counter(), Counter()
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    c[op] += int(r[iex]); s[op] += int(r[ismp]) if r[ismp].isdigit() else 0
tot = sum(c.values()); ts = sum(s.values())
print("  total warp instructions", tot)
for op, n in c.most_common(16):
    print("  %-10s %6.2f%% of instructions  %5.1f%% of stall samples" % (op, 100 * n / tot, 100 * s[op] / max(ts, 1)))
