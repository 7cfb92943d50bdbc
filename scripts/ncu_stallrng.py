"""Per source-line range: executed instructions, stall samples and the main stall reasons.
usage: ncu_stallrng.py file.csv srcname lo-hi[:label] ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
rngs = []
for a in sys.argv[3:]:
    r, _, lab = a.partition(':')
    lo, hi = r.split('-')
    rngs.append((int(lo), int(hi), lab or r))
cols = ['stall_no_inst', 'stall_wait', 'stall_long_sb', 'stall_short_sb', 'stall_branch_resolving', 'stall_barrier', 'stall_math', 'stall_not_selected', 'stall_selected']
hdr = None; fname = ''; per = {}
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] != '-': continue
    try:
        vals = [int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('Warp Stall Sampling (All Samples)')])] + [int(r[hdr.index(c)] or 0) for c in cols]
    except ValueError: continue
    key = int(r[0]) if fname == want else -1
    old = per.get(key, [0] * len(vals)); per[key] = [a + b for a, b in zip(old, vals)]
tot = [sum(v[i] for v in per.values()) for i in range(2 + len(cols))]
print('%-22s %6s %6s | ' % ('range', 'instr%', 'stall%') + ' '.join('%8s' % c[6:14] for c in cols))
for lo, hi, lab in rngs + [(-1, -1, 'other files')]:
    s = [sum(v[i] for k, v in per.items() if lo <= k <= hi) for i in range(2 + len(cols))]
    print('%-22s %6.1f %6.1f | ' % (lab, 100 * s[0] / tot[0], 100 * s[1] / tot[1]) + ' '.join('%8.1f' % (100 * s[2 + i] / tot[1]) for i in range(len(cols))))
print('%-22s %6s %6s | ' % ('all', '', '') + ' '.join('%8.1f' % (100 * tot[2 + i] / tot[1]) for i in range(len(cols))))
