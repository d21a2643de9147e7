"""Group per-line instruction counts of a profile into code regions (function ranges of the source files)."""
import csv, re, sys, collections
sass_csv, lines_txt = sys.argv[1], sys.argv[2]
updates = float(sys.argv[3]) if len(sys.argv) > 3 else None
def regions(path):
    out = []; name = None
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'(?:template.*>\s*)?(?:__device__|__global__|static|inline|struct)\b.*?\b(\w+)\s*(\(|\{|$)', l)
        if l.startswith('static '): l = l[7:]
        if l.startswith('__device__') or l.startswith('__global__') or l.startswith('struct ') or l.startswith('template'):
            m2 = re.search(r'(\w+)\s*\(', l) or re.search(r'struct\s+(\w+)', l)
            if m2 and not l.startswith('template'): out.append((i, m2.group(1)))
    return out
reg = {f: regions('gibbssampling_b200/csrc/' + f) for f in ('gibbs_device.cuh', 'gibbs_kernels.cuh', 'gibbs_drift_dev.cuh', 'gibbs_motif.cuh', 'gibbs_cluster.cuh')}
def region_of(f, l):
    if f not in reg: return f
    name = '?'
    for start, n in reg[f]:
        if start <= l: name = n
        else: break
    return f"{f.split('_')[1][:3]}:{name}"
off2line = {}; cur = None; started = False
for ln in open(lines_txt):
    if ln.startswith('//--------------------- .text'):
        if started: break
        started = True; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(\S+)', ln)
    if m: off2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(sass_csv))); hdr = rows[1]
ia, ismp, iinst = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
base = int(rows[2][ia], 16)
by = collections.defaultdict(lambda: [0, 0]); ops = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    (key, op) = off2line.get(int(r[ia], 16) - base, (('?', 0), '?'))
    n = int(r[iinst]); s = int(r[ismp])
    g = region_of(*key); by[g][0] += n; by[g][1] += s; tot += n
    ops[op.split('.')[0]] += n
tots = sum(v[1] for v in by.values())
for g, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][0])[:25]:
    per = f"{n/updates:7.1f}/upd" if updates else ""
    print(f"{100*n/tot:5.1f}% inst {100*s/tots:5.1f}% smp {per}  {g}")
print("opcodes:", ", ".join(f"{o}:{100*c/tot:.1f}%" for o, c in ops.most_common(22)))
