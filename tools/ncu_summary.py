"""Summarises an .ncu-rep (read here, no GPU needed) into the few lines DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_tex_mem_texture.sum", "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_texture.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]
for row in rows[2:]:
    d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {u[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    tot = 0; recs = []
    stall = collections.Counter()
    for r in rows[2:]:
        try:
            s = int(r[ix["# Samples"]])
        except Exception:
            continue
        tot += s
        recs.append((s, r[ix["Source"]].strip()[:70], r[ix["Instructions Executed"]]))
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stall[h] += int(r[ix[h]])
                except Exception:
                    pass
    print(f"warp-state samples: {tot}")
    print("  by reason:", ", ".join(f"{k}={v} ({100.0 * v / max(tot, 1):.1f}%)" for k, v in stall.most_common(8)))
    recs.sort(reverse=True)
    print("  top instructions by samples:")
    for s, t, n in recs[:12]:
        print(f"    {s:7d} ({100.0 * s / max(tot, 1):4.1f}%)  x{n:>10s}  {t}")
