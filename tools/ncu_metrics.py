"""key metrics of the first kernel of an `ncu --set full` report:  ncu -i X.ncu-rep --page raw --csv | python tools/ncu_metrics.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct"]
seen = set()
for w in want:
    for i, h in enumerate(hdr):
        if h == w and h not in seen:
            seen.add(h)
            print("| `%s` | %s %s |" % (h, vals[i], units[i]))
for i, h in enumerate(hdr):
    if ("pipe" in h and "pct" in h) and h not in seen:
        try:
            if float(vals[i].replace(",", "")) >= 3:
                print("| `%s` | %s %s |" % (h, vals[i], units[i]))
        except ValueError:
            pass
