"""Per-SASS-instruction executed counts of one kernel from an .ncu-rep (source page).
usage: python profiles/ncu_hot.py rep kernel-regex [min_pct]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if len(r) > 5 and r[0] == 'Address':
        if hdr is not None: break   # first kernel instance only
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ia, it, isrc, isamp = hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('Source'), hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data); stot = sum(int(r[isamp]) for r in data)
print('total warp-instructions', tot, 'samples', stot, 'sass lines', len(data))
for n, r in enumerate(data):
    v = int(r[ia]); s = int(r[isamp])
    if 100.0 * v / tot >= minpct or 100.0 * s / max(stot, 1) >= minpct:
        print('%4d exec=%5.2f%% thr=%5s stall-samples=%5.2f%%  %s' % (n, 100.0 * v / tot, r[it], 100.0 * s / max(stot, 1), r[isrc].strip()[:100]))
