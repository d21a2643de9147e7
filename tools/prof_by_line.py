"""Join an ncu SASS source page (csv) with nvdisasm -g line info: instructions and stall samples per CUDA line.
usage: prof_by_line.py <ncu_sass.csv> <nvdisasm_lines.txt> [top]"""
import csv, re, sys, collections, glob
sass_csv, lines_txt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
off2line = {}
cur = None
started = False
for ln in open(lines_txt):
    if ln.startswith('//--------------------- .text'):
        if started: break
        started = True
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*)', ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ismp, iinst = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(rows[2][ia], 16)
by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    off = int(r[ia], 16) - base
    key = off2line.get(off, ('?', 0))
    n, s = int(r[iinst]), int(r[ismp])
    by[key][0] += n; by[key][1] += s
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: by[key][2][hdr[i]] += v
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i}, samples {tot_s}")
src_cache = {}
def src(f, l):
    if f not in src_cache:
        c = glob.glob(f'gibbssampling_b200/csrc/{f}')
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[l-1].strip()[:90] if 0 < l <= len(L) else ''
for key, (n, s, st) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    top_st = ','.join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    print(f"{100*s/tot_s:5.1f}% smp {100*n/tot_i:5.1f}% inst  {key[0]}:{key[1]:<4} [{top_st}]  {src(*key)}")
