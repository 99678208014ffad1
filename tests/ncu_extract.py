"""Compact an `ncu -i X.ncu-rep --page raw --csv` dump to the metrics DESIGN.md / bench.py cite:
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tests/ncu_extract.py raw.csv profiles/out.csv"""
import csv, sys
WANT = ["ID", "Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(w) for w in WANT if w in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([(r[i].split("(")[0] if hdr[i] == "Kernel Name" else r[i]) for i in idx])
print("wrote", sys.argv[2], len(data), "launches")
