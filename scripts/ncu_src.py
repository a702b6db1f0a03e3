"""Summarise `ncu --page source --csv --print-source sass` output: op mix, loop regions, smem excess."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
ia = h.index('Address'); isrc = h.index('Source'); ie = h.index('Instructions Executed')
iss = h.index('Warp Stall Sampling (All Samples)')
iw = h.index('L1 Wavefronts Shared'); iwi = h.index('L1 Wavefronts Shared Ideal')
data = [r for r in rows[hi + 1:] if len(r) > max(ie, iss, iw, iwi) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data); tots = max(1, sum(int(r[iss]) for r in data))
print('kernel', rows[0][1][:80] if rows[0] else '')
print('total warp-inst', tot, 'samples', tots, 'static', len(data))
byop = collections.Counter(); bys = collections.Counter()
def opof(s):
    t = s.split()
    op = t[1] if t[0].startswith('@') else t[0]
    return op.split('.')[0]
for r in data:
    byop[opof(r[isrc])] += int(r[ie]); bys[opof(r[isrc])] += int(r[iss])
for op, c in byop.most_common(18):
    print(f'{op:10s} {c:12d} {100*c/tot:5.1f}%  samples {100*bys[op]/tots:5.1f}%')
prev = None; start = 0; segs = []
for i, r in enumerate(data):
    e = int(r[ie])
    if prev is None or abs(e - prev) > 0.02 * max(e, prev, 1):
        if prev is not None: segs.append((start, i - 1, prev))
        start = i
    prev = e
segs.append((start, len(data) - 1, prev))
for s in segs:
    n = s[1] - s[0] + 1
    smp = sum(int(data[i][iss]) for i in range(s[0], s[1] + 1))
    if n * s[2] > 0.01 * tot or smp > 0.02 * tots:
        ops = collections.Counter(opof(data[i][isrc]) for i in range(s[0], s[1] + 1))
        print(f'inst[{s[0]:5d}-{s[1]:5d}] n={n:5d} exec={s[2]:10d} share={100*n*s[2]/tot:5.1f}% samples={100*smp/tots:5.1f}%  {dict(ops.most_common(5))}')
wf = sum(int(r[iw]) for r in data); wfi = sum(int(r[iwi]) for r in data)
print('shared wavefronts', wf, 'ideal', wfi)
for r in data:
    if int(r[iw]) > 0 and int(r[iw]) > 1.3 * int(r[iwi]) and int(r[iw]) > 0.01 * wf:
        print('  ', r[isrc][:70], r[iw], r[iwi])
# top stall instructions
top = sorted(data, key=lambda r: -int(r[iss]))[:12]
print('top stall samples:')
for r in top: print('  ', r[iss].rjust(6), r[ie].rjust(10), r[isrc][:80])
