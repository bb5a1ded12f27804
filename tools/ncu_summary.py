"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per launch:  python ncu_summary.py raw.csv > summary.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_read"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")]
idx = [(hdr.index(k) if k in hdr else None, n) for k, n in want]
w = csv.writer(sys.stdout)
w.writerow([n + (f" [{units[i]}]" if i is not None and units[i] else "") for i, n in idx])
for r in rows[2:]:
    out = []
    for i, n in idx:
        v = r[i] if i is not None else ""
        if n == "kernel":
            v = v.replace("CUtensorMap_st, ", "").replace("(int)", "").replace("(bool)", "")[:120]
        out.append(v)
    w.writerow(out)
