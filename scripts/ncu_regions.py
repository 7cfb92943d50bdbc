"""Executed warp instructions of a kernel split at its block barriers (SASS order), with the FP64 share of each region.
Input: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; sass = {}
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] in ('-', ''):
        continue
    ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); st = hdr.index('Warp Stall Sampling (All Samples)')
    try:
        sass[int(r[2], 16)] = (r[3], int(r[ie]), int(r[te]), int(r[st]), r[0])
    except ValueError:
        pass
tot = sum(v[1] for v in sass.values()); stot = sum(v[3] for v in sass.values()) or 1
print('SASS instructions', len(sass), 'executed warp instr', tot)
reg = []; cur = [0, 0, 0, 0, None, None]
for a in sorted(sass):
    op, n, tn, s, line = sass[a]
    if cur[4] is None: cur[4] = line
    cur[0] += n; cur[2] += tn; cur[3] += s
    mn = op.split()[0] if not op.startswith('@') else op.split()[1]
    if mn.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU.RSQ64', 'MUFU.RCP64')): cur[1] += n
    if 'BAR.SYNC' in op or 'EXIT' in op or 'RET' in op:
        cur[5] = line; reg.append(cur); cur = [0, 0, 0, 0, None, None]
reg.append(cur)
for c in reg:
    if c[0]: print('lines %5s..%5s  %5.1f%% instr  %5.1f%% stall  fp64 share %4.1f%%  lanes %4.1f' % (c[4], c[5], 100 * c[0] / tot, 100 * c[3] / stot, 100 * c[1] / c[0], c[2] / c[0]))
