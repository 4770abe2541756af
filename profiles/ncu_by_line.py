"""Per-source-line totals (instructions executed, stall samples) of one kernel: joins `ncu --page source --csv` (SASS view, one
row per instruction in address order) with `nvdisasm -c --print-line-info` of the same cubin (same order).
   python profiles/ncu_by_line.py <ncu_source.csv> <nvdisasm_lineinfo.txt> <kernel name substring> [top]"""
import csv
import re
import sys

src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ins = []
for r in rows[2:]:
    try:
        ins.append((r[i_src].strip(), int(r[i_ex]), int(r[i_s])))
    except (ValueError, IndexError):
        pass
# nvdisasm: walk the kernel's section, remember the current "//## File "...", line N" marker for every instruction
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('\t.section') and kname in l and '.text.' in l)
cur = ("?", 0)
locs = []
for l in lines[start + 1:]:
    if l.startswith('\t.section') or l.startswith('//---------------------'):
        if locs:
            break
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        locs.append(cur)
print("instructions: ncu %d, nvdisasm %d" % (len(ins), len(locs)))
n = min(len(ins), len(locs))
agg = {}
for (sass, ex, sm), loc in zip(ins[:n], locs[:n]):
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += ex; a[1] += sm; a[2] += 1
tot_ex = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total executed %d, samples %d" % (tot_ex, tot_s))
srcs = {}
def text(loc):
    f, ln = loc
    if f not in srcs:
        try:
            import glob
            path = glob.glob('spatio_temporal_voxel_layer_b200/csrc/' + f)[0]
            srcs[f] = open(path).read().splitlines()
        except Exception:
            srcs[f] = []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
print("-- by stall samples")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samples %5.1f%% instr  %s:%d  %s" % (100.0 * a[1] / max(tot_s, 1), 100.0 * a[0] / max(tot_ex, 1), loc[0], loc[1], text(loc)))
print("-- by instructions executed")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * a[0] / max(tot_ex, 1), 100.0 * a[1] / max(tot_s, 1), loc[0], loc[1], text(loc)))
