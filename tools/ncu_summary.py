"""Summarise an .ncu-rep: key raw metrics + stall samples aggregated per source window.  python -m tools.ncu_summary rep [window]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h}\t{u}\t{v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, iex, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [(r[ia].strip(), int(r[iex] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > iex]
ts = sum(d[2] for d in data)
print("total samples", ts, "instructions", len(data))
for i in range(0, len(data), win):
    w = data[i:i + win]
    sm = sum(d[2] for d in w)
    if sm < 0.01 * ts:
        continue
    ops = collections.Counter(re.sub(r"^(@!?U?P\d+\s+)", "", d[0]).split(" ")[0] for d in w)
    hot = max(w, key=lambda d: d[2])
    print(f"{i:5d} {100 * sm / ts:5.1f}%  exec~{max(d[1] for d in w):9d} {dict(ops.most_common(5))}  hottest: {hot[0][:60]} ({100 * hot[2] / ts:.1f}%)")
