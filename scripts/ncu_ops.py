"""Aggregate an `ncu -i X.ncu-rep --page source --csv` (SASS) export by opcode: share of stall samples, executed count, cycles/instr.
usage: python scripts/ncu_ops.py sass.csv [warp_iterations]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[1]
iS = h.index('# Samples'); iE = h.index('Instructions Executed'); iSrc = h.index('Source')
st = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = 0
for r in rows[2:]:
    try:
        s = int(r[iS]); e = int(r[iE])
    except ValueError:
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iSrc])
    if not m:
        continue
    a = agg[m.group(2)]
    a[0] += s; a[1] += e; a[2] += 1
    for i, c in st:
        if r[i].isdigit():
            a[3][c[6:]] += int(r[i])
    tot += s
print('op        samples%%  exec/unit  static  top stalls   (unit = %g)' % norm)
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:26]:
    print('%-10s %5.1f%% %9.1f %6d  %s' % (op, 100 * a[0] / tot, a[1] / norm, a[2], ' '.join('%s=%d' % kv for kv in a[3].most_common(3))))
print('total samples', tot, ' executed/unit', sum(a[1] for a in agg.values()) / norm)
