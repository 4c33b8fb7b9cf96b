"""Top stall sites from `ncu -i rep --page source --csv --kernel-id :::N` output.  python scripts/ncu_src_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
ia, isrc, ins, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
cols = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_not_selected', 'stall_selected', 'stall_branch_resolving',
        'stall_math', 'stall_dispatch', 'stall_no_inst', 'stall_mio', 'stall_barrier', 'stall_lg', 'stall_membar']
ci = [h.index(c) for c in cols]
data = [r for r in rows[2:] if len(r) > iex and r[ins].isdigit()]
tot = sum(int(r[ins]) for r in data)
totex = sum(int(r[iex]) for r in data)
print("total samples", tot, "warp instructions executed", totex)
for c, i in zip(cols, ci):
    print(f"  {c:24s} {sum(int(r[i] or 0) for r in data):7d}")
base = int(data[0][ia], 16)
for r in sorted(data, key=lambda r: -int(r[ins]))[:n]:
    st = ' '.join(f"{c[6:10]}={r[i]}" for c, i in zip(cols, ci) if r[i] not in ('0', ''))
    print(f"{int(r[ia], 16) - base:6x} {int(r[ins]):5d} ex={r[iex]:>7s} {r[isrc].strip()[:58]:58s} {st}")
