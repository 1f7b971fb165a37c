"""Summarise an `ncu --page source --csv` export: stall reasons overall and the hottest SASS
instructions.   python tools/ncu_stalls.py gpurun_out/k3src.csv [n_top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError):
        return 0.0


tot = sum(f(r, '# Samples') for r in data)
print('total samples', tot, 'instructions', len(data))
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {k: sum(f(r, k) for r in data) for k in reasons}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print('  %-24s %6.2f%%' % (k, 100 * v / tot))
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:ntop]:
    rs = sorted(reasons, key=lambda k: -f(r, k))[:2]
    print('%s %6.2f%% %-72s %s' % (r[ix['Address']][-6:], 100 * f(r, '# Samples') / tot, r[ix['Source']][:72],
                                 ' '.join('%s=%d' % (k[6:], f(r, k)) for k in rs)))
