"""Print a short per-kernel digest of an .ncu-rep (`ncu --set full`): duration, tensor pipe, L2->SM fabric, DRAM, stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in txt.splitlines() if not l.startswith("==")))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("-" * 100)
    for k in KEYS:
        cands = [h for h in hdr if h.endswith(k)]
        for h in cands[:1]:
            print(f"{k:82s} {units[hdr.index(h)]:10s} {d[h]}")
    stalls = [(float(d[h] or 0), h) for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d[h]]
    for v, h in sorted(stalls, reverse=True)[:6]:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:.2f}")
