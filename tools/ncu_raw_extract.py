"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the short text extracts kept under profiles/.
usage: python tools/ncu_raw_extract.py raw.csv "header comment" > profiles/NAME.txt"""
import csv, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__cluster_max_active", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none"))
for r in rows[2:]:
    print(f"\n== {r[H.index('Kernel Name')]}")
    for k in KEEP:
        if k in H:
            print(f"{k} = {r[H.index(k)]} {U[H.index(k)]}")
    stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(H) if "issue_stalled" in h and "per_issue_active" in h), reverse=True)[:8]
    print("top stalls (warps per issue-active cycle): " + ", ".join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in stalls))
