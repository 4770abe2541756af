"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` (first pass only)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
names = ['stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_wait', 'stall_branch_resolving', 'stall_math', 'stall_lg',
         'stall_mio', 'stall_membar', 'stall_no_inst', 'stall_dispatch', 'stall_not_selected', 'stall_selected', 'stall_sleep']
idx = [hdr.index(n) for n in names]
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((int(r[i_s]), n, r))
    except (ValueError, IndexError):
        continue
data = data[:len(data) // 2]
print("total samples", sum(d[0] for d in data), "instructions", len(data))
agg = {}
for s, n, r in data:
    for nm, i in zip(names, idx):
        try:
            agg[nm] = agg.get(nm, 0) + int(r[i])
        except ValueError:
            pass
print({k: v for k, v in agg.items() if v})
for s, n, r in sorted(data, reverse=True)[:top]:
    print("%6d  #%4d ex=%8s %s  %s" % (s, n, r[i_ex], " ".join("%s=%s" % (nm[6:], r[i]) for nm, i in zip(names, idx) if r[i] not in ('0', '')), r[i_src][:90]))
