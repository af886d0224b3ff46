"""Export the key metrics of `ncu --set full` captures (.ncu-rep, read with `ncu -i … --page raw --csv`) into one CSV.
usage: python tools/export_ncu_full.py out.csv rep1.ncu-rep [rep2 ...]"""
import csv, subprocess, sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes.sum.per_second"]
out = csv.writer(open(sys.argv[1], "w"))
first = True
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    if first:
        out.writerow([hdr[i] for i in idx]); out.writerow([units[i] for i in idx]); first = False
    for r in rows[2:]:
        out.writerow([r[i] for i in idx])
