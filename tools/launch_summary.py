"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
the launch sequence of one step.  usage: launch_summary.py file.csv [launches_per_step] [--seq]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size'); ui = h.index('Metric Unit')
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    t = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': t /= 1000.0
    elif r[ui] == 'ms': t *= 1000.0
    seq.append((r[ki], r[gi], t))
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else len(seq)
seq = seq[:n]
def short(nm):
    base = nm.split('<')[0].split('(')[0].replace('void ', '')
    tpl = nm[len(nm.split('<')[0]):][:48] if '<' in nm else ''
    return (base + tpl)[:80]
agg = {}
for nm, g, t in seq:
    a = agg.setdefault(short(nm), [0.0, 0]); a[0] += t; a[1] += 1
tot = sum(t for _, _, t in seq)
print("launches %d  total %.1f us" % (len(seq), tot))
for nm, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%9.1f us %5.1f%% %4d  %s" % (t, 100 * t / tot, c, nm))
if '--seq' in sys.argv:
    for nm, g, t in seq: print("%8.1f  %-18s %s" % (t, g, short(nm)))
