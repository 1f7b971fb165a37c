"""Join an `ncu --page source --csv` export (per-SASS-instruction stall samples) with
`nvdisasm --print-line-info` of the same cubin and aggregate the samples by device function and
by CUDA source line.
    python tools/ncu_sass_lines.py k3src.csv k3.disasm <kernel substring> [n_top]"""
import csv
import re
import sys

src_csv, disasm, kern = sys.argv[1], sys.argv[2], sys.argv[3]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 50

# --- disassembly: offset -> (function label, file, line) for the chosen kernel section
info = {}
in_sec, func, fl = False, None, ('?', 0)
for ln in open(disasm):
    if ln.startswith('\t.section\t.text.') or ln.startswith('.section'):
        in_sec = kern in ln
        continue
    if not in_sec:
        continue
    m = re.match(r'^(\$?_Z\S+):', ln)
    if m:
        name = m.group(1)
        mm = re.findall(r'\d+([a-z_]+[a-z0-9_]*)', name.split('$')[-1])
        func = next((x for x in mm if x in ('offdiag_tile', 'diag_factor_blk', 'forward_row_w', 'k3_factor_flow', 'k2_dofmap_assemble', 'k4_solve_recover')), name[-40:])
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        fl = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        info[int(m.group(1), 16)] = (func, fl[0], fl[1], m.group(2).strip())

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix['Address']], 16)
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError):
        return 0.0


tot = sum(f(r, '# Samples') for r in data)
byfunc, byline = {}, {}
for r in data:
    off = int(r[ix['Address']], 16) - base
    fn, fi, li, sass = info.get(off, ('?', '?', 0, ''))
    n = f(r, '# Samples')
    byfunc[fn] = byfunc.get(fn, 0) + n
    key = (fn, fi, li)
    e = byline.setdefault(key, [0.0, {}, 0.0])
    e[0] += n
    e[2] += f(r, 'Instructions Executed')
    for k in reasons:
        e[1][k] = e[1].get(k, 0) + f(r, k)
print('total samples', tot)
for fn, n in sorted(byfunc.items(), key=lambda kv: -kv[1]):
    print('  %-20s %6.2f%%' % (fn, 100 * n / tot))
for key, (n, rs, ie) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:ntop]:
    top = sorted(rs.items(), key=lambda kv: -kv[1])[:3]
    print('%5.2f%% %-16s %-22s %4d  inst %-10d %s' % (100 * n / tot, key[0], key[1], key[2], ie,
                                                      ' '.join('%s=%.1f%%' % (k[6:], 100 * v / tot) for k, v in top)))
