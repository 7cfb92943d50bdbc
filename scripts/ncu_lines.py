"""Top CUDA source lines by executed warp instructions / stall samples.
Input: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
out = []; fname = ''
hdr = None
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] != '-':
        continue
    ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); st = hdr.index('Warp Stall Sampling (All Samples)')
    try:
        out.append((int(r[ie]), int(r[te]), int(r[st]), fname, r[0], r[1][:105]))
    except ValueError:
        pass
tot = sum(o[0] for o in out) or 1; stot = sum(o[2] for o in out) or 1
print('total warp instructions', tot, ' stall samples', stot)
key = 2 if len(sys.argv) > 3 and sys.argv[3] == 'stall' else 0
for o in sorted(out, key=lambda o: -o[key])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%5.1f%% instr %5.1f%% stall lanes %4.1f | %s:%s %s' % (100 * o[0] / tot, 100 * o[2] / stot, o[1] / max(o[0], 1), o[3], o[4], o[5]))
