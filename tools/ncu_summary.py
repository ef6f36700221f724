"""`ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py out.csv`: keep the metrics the roofline
discussion needs, one row per profiled launch (proper CSV quoting: kernel names contain commas)."""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if h in WANT]
with open(sys.argv[1], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([f"{hdr[i]} [{units[i]}]" for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
