"""Condense .ncu-rep files into the metric lines DESIGN.md quotes.  Usage: ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys
KEEP = ("launch__block_size", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_tc", "sm__pipe_tensor", "sm__warps_active.avg.pct",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "smsp__warps_active.avg.per_cycle_active", "sm__throughput.avg.pct", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_uniform", "launch__shared_mem_per_block_dynamic")
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("=====", rep, "(no data)"); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("=====", rep)
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name": print(" ", v[:110])
        if any(h.startswith(k) for k in KEEP):
            try:
                if float(v.replace(",", "")) == 0 and "stalled" in h: continue
            except ValueError:
                pass
            print(f"  {h} [{u}] = {v}")
