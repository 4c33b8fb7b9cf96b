"""Reduce `ncu -i rep --page raw --csv` output to the columns quoted in DESIGN.md / profiles/README.md; keeps the LAST
launch of every distinct kernel name + grid (warm).  python scripts/ncu_summarize.py raw.csv out.csv"""
import csv
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum"]
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
idx = [h.index(c) for c in COLS if c in h]
stalls = [i for i, c in enumerate(h) if "issue_stalled" in c and c.endswith("per_issue_active.ratio") and "not_issued" not in c]
last = {}
for r in rows[2:]:
    if len(r) < len(h):
        continue
    last[(r[h.index("Kernel Name")], r[h.index("launch__grid_size")])] = r
w = csv.writer(open(sys.argv[2], "w"))
w.writerow([h[i] for i in idx] + ["top_stalls_per_issue"])
w.writerow([units[i] for i in idx] + [""])
for r in last.values():
    st = sorted(((float(r[i].replace(",", "") or 0), h[i].split("issue_stalled_")[1].split("_per_issue")[0]) for i in stalls), reverse=True)[:4]
    w.writerow([r[i] for i in idx] + [" ".join(f"{n}={v:.2f}" for v, n in st)])
print(f"{len(last)} kernels -> {sys.argv[2]}")
