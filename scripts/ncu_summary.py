"""Text summary of an .ncu-rep (raw page + per-instruction page): the metrics DESIGN.md quotes.  usage: ncu_summary.py rep [units_per_launch]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, un, vals = rows[0], rows[1], rows[2]
g = lambda k: vals[hdr.index(k)] if k in hdr else "n/a"
print("kernel:", g("Kernel Name"), " grid", g("launch__grid_size"), "x", g("launch__block_size"), "threads,", g("launch__registers_per_thread"), "regs,",
      g("launch__shared_mem_per_block"), "KB smem/block")
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    if k in hdr:
        print(f"  {k} = {g(k)} {un[hdr.index(k)]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and float(vals[i] or 0) >= 0.3:
        print(f"  {h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} = {float(vals[i]):.2f} warps per issue")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]; data = rows[2:]
    isrc, iex, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[iex]) for r in data)
    print(f"  warp instructions executed: {tot}" + (f" = {tot / units:.1f} per unit ({units:.0f} units per launch)" if units else ""))
    ops = collections.Counter()
    for r in data:
        t = r[isrc].strip().split()
        o = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
        ops[o.split(".")[0]] += int(r[iex])
    print("  instruction mix (share of executed):", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(12)))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:8]
    ts = sum(int(r[isamp]) for r in data)
    print("  hottest instructions by stall samples:")
    for i in top:
        print(f"    {100 * int(data[i][isamp]) / max(ts, 1):5.1f}%  {data[i][isrc].strip()[:70]}")
