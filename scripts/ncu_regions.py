"""Stall samples of k_fit by code region (objective / optimiser / called helpers) from an ncu report and the object
that was profiled:  python scripts/ncu_regions.py REPORT.ncu-rep OBJECT.o"""
import csv, io, os, re, subprocess, sys, tempfile

rep, obj = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; sass = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): sass.append(dict(zip(hdr, r)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = ('?', 0); lines = []; func = 'main'; funcs = []
for l in dis:
    if l.startswith('.text.'):
        inside = '_Z5k_fitPKdx' in l; func = 'k_fit'; continue
    if not inside: continue
    m = re.match(r'\$_Z5k_fitPKdx[^$]*\$(.*):', l) or re.match(r'\$(__internal[^:]*):', l)
    if m: func = m.group(1)[:60]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l): lines.append(cur); funcs.append(func)
print(len(sass), len(lines))
keys = ['stall_wait', 'stall_no_inst', 'stall_selected', 'stall_short_sb', 'stall_branch_resolving', 'stall_long_sb', 'stall_math', 'stall_not_selected', 'stall_dispatch', 'stall_lg', 'stall_mio']
def region(k):
    f, ln = lines[k]; fn = funcs[k]
    if fn != 'k_fit':
        if 'fit_eval' in fn: return 'objective(call)'
        return fn[-14:]
    if f in ('fit_eval.h', 'exp_glibc.h', 'sm_32_intrinsics.hpp'): return 'objective'
    if f == 'sxs_exact.cu' and 150 <= ln <= 330: return 'objective'
    return 'kfit:' + f
agg = {}
for k, d in enumerate(sass[:len(lines)]):
    a = agg.setdefault(region(k), [0] * (len(keys) + 3))
    a[0] += float(d['# Samples'] or 0); a[1] += float(d['Instructions Executed'] or 0); a[2] += float(d['Thread Instructions Executed'] or 0)
    for i, kk in enumerate(keys): a[3 + i] += float(d[kk] or 0)
ts = sum(a[0] for a in agg.values())
print('%-18s %6s %6s %5s ' % ('region', 'samp%', 'Ginst', 'lanes') + ' '.join('%8s' % k[6:14] for k in keys))
for r, a in sorted(agg.items(), key=lambda z: -z[1][0]):
    print('%-18s %6.1f %6.2f %5.1f ' % (r, 100 * a[0] / ts, a[1] / 1e9, a[2] / max(a[1], 1)) + ' '.join('%8.1f' % (100 * x / ts) for x in a[3:]))
