"""Summarise ncu captures: python scripts/ncu_summary.py TAG  (reads gpurun_out/TAG_*)."""
import csv, subprocess, sys
tag = sys.argv[1]
keys = ["gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","launch__registers_per_thread","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers","launch__shared_mem_per_block_dynamic","sm__throughput.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__inst_executed.sum","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
out = []
for f in ["lauum", "step"]:
  txt = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_{f}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(txt.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  out.append(f"## {f}: " + " | ".join(r[hdr.index('Kernel Name')][:40] + " grid " + r[hdr.index('Grid Size')] for r in data))
  for k in keys:
    if k in hdr:
      i = hdr.index(k)
      out.append(f"- {k} [{units[i]}]: " + " | ".join(r[i] for r in data))
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_launches.csv")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii, gi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
d = {}
for r in rows[1:]:
  d.setdefault((int(r[ii]), r[ki].split("(")[0][-24:], r[gi]), {})[r[mi]] = r[vi]
out.append("## launch list (one hb_nll_grad_batched call; cold-cache, serialised)")
out.append("| # | kernel | grid | time us | tensor % | fp64 % | warps % | DRAM rd MB | DRAM wr MB |")
out.append("|---|---|---|---|---|---|---|---|---|")
tot = 0.0
def num(x): return float(x.replace(",", ""))
for k in sorted(d):
  m = d[k]
  t = num(m["gpu__time_duration.sum"]); tot += t
  scale = 1e-3 if t > 1e3 else 1.0
  rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
  out.append(f"| {k[0]} | {k[1]} | {k[2]} | {t*scale:.1f} | " + " | ".join(f"{num(m[x]):.1f}" for x in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active")) + f" | {rd} | {wr} |")
print("\n".join(out))
