"""Per-SOURCE-line stall samples and executed instructions of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on).  usage: python profiles/ncu_lines.py rep kernel-regex [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data, nk = None, None, [], 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) == 2 and r[0] == 'Function Name':
        continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        data.append((cur, r))
isamp, iexe = hdr.index('# Samples'), hdr.index('Instructions Executed')
ilsb = hdr.index('stall_long_sb')
tot = sum(int(r[isamp]) for _, r in data); etot = sum(int(r[iexe]) for _, r in data)
print('samples', tot, 'warp-instructions', etot)
for f, r in sorted(data, key=lambda fr: -int(fr[1][isamp]))[:top]:
    print('%5.2f%% samples %5.2f%% exec long_sb %5.2f%%  %s:%s  %s' % (100.0 * int(r[isamp]) / tot, 100.0 * int(r[iexe]) / etot,
          100.0 * int(r[ilsb]) / tot, f, r[0], r[1].strip()[:110]))
