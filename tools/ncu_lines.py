"""Aggregate an `ncu --page source --csv --print-source cuda` export by CUDA source line:
share of warp-stall samples per line with its two dominant stall reasons.
    python tools/ncu_lines.py gpurun_out/k3cuda.csv [n_top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur, hdr, out = None, None, []
for r in rows:
    if r and r[0] == 'File Name':
        cur, hdr = r[1].split('/')[-1], None
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            n = float(d.get('# Samples', '0') or 0)
        except ValueError:
            n = 0
        if n > 0:
            out.append((n, cur, int(d['Line No']), d['Source'].strip()[:90], d))
tot = sum(o[0] for o in out)
print('total samples', tot)
byfile = {}
for n, f, l, s, d in out:
    byfile[f] = byfile.get(f, 0) + n
for f, n in sorted(byfile.items(), key=lambda kv: -kv[1]):
    print('  %-28s %6.2f%%' % (f, 100 * n / tot))
for n, f, l, s, d in sorted(out, key=lambda x: -x[0])[:ntop]:
    rs = sorted([(float(d[k] or 0), k) for k in d if k.startswith('stall_') and 'Not' not in k], reverse=True)[:2]
    print('%5.2f%% %-22s %4d %-90s %s' % (100 * n / tot, f, l, s, ' '.join('%s=%.0f' % (k[6:], v) for v, k in rs)))
