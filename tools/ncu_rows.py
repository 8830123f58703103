#!/usr/bin/env python
"""Print selected metrics of every launch in an `ncu -i X.ncu-rep --page raw --csv` dump as a markdown table.
Usage: ncu -i rep --page raw --csv | python tools/ncu_rows.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma subpipe % (active)"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem tensor-core wavefronts %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "L2->SM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__cycles_elapsed.avg", "cycles"),
        ("launch__registers_per_thread", "regs")]
cols = []
for key, label in WANT:
    for i, h in enumerate(hdr):
        if h == key:
            cols.append((i, label)); break
print("| launch | kernel | " + " | ".join(l for _, l in cols) + " |")
print("|---|---|" + "---:|" * len(cols))
ki = hdr.index("Kernel Name")
for r in data:
    vals = []
    for i, _ in cols:
        v, u = r[i], units[i]
        vals.append(f"{v} {u}".strip())
    print(f"| {r[0]} | `{r[ki][:28]}` | " + " | ".join(vals) + " |")
if "--list" in sys.argv:
    for h in hdr:
        print(h)
