"""Executed warp instructions / stall samples per source-line range of one file.
usage: ncu_ranges.py file.csv srcname lo-hi[:label] ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
rngs = []
for a in sys.argv[3:]:
    r, _, lab = a.partition(':')
    lo, hi = r.split('-')
    rngs.append((int(lo), int(hi), lab or r))
hdr = None; fname = ''; per = {}; tot = 0; stot = 0
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] != '-': continue
    ie = hdr.index('Instructions Executed'); st = hdr.index('Warp Stall Sampling (All Samples)')
    try: n, s = int(r[ie]), int(r[st])
    except ValueError: continue
    tot += n; stot += s
    if fname == want: per[int(r[0])] = (n, s)
    else: per.setdefault(-1, [0, 0]); per[-1] = (per[-1][0] + n, per[-1][1] + s)
print('total', tot)
acc = 0
for lo, hi, lab in rngs:
    n = sum(v[0] for k, v in per.items() if lo <= k <= hi); s = sum(v[1] for k, v in per.items() if lo <= k <= hi)
    acc += n
    print('%-28s %5.1f%% instr %5.1f%% stall' % (lab, 100 * n / tot, 100 * s / stot))
print('%-28s %5.1f%% instr' % ('other files (intrinsics)', 100 * per.get(-1, (0, 0))[0] / tot))
print('%-28s %5.1f%% instr' % ('unlisted', 100 * (tot - acc - per.get(-1, (0, 0))[0]) / tot))
