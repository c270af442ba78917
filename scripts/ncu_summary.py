"""Summarise an .ncu-rep (raw page) into the metrics DESIGN.md / profiles/ quote.  usage: ncu_summary.py file.ncu-rep [launch-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
stall = [n for n in h if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_warp_active.pct") is False]
stall2 = [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith(".ratio")]
for d in data:
    print("=" * 100)
    for w in WANT:
        if w in h:
            i = h.index(w); print(f"{w:90s} {units[i]:14s} {d[i]}")
    st = []
    for n in stall2:
        i = h.index(n)
        try: st.append((float(d[i].replace(",", "")), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError: pass
    st.sort(reverse=True)
    print("top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:6]))
