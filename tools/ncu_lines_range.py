"""Like ncu_lines.py but prints samples per source line in source order for a line range.
usage: ncu_lines_range.py <ncu_source.csv> <nvdisasm.txt> <kernel-substring> <file.cu> <first> <last>"""
import collections, csv, re, sys
src_csv, dis, kname, cu, l0, l1 = sys.argv[1:7]
l0, l1 = int(l0), int(l1)
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l][0]
inst, pend = [], None
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if pend is None:
            pend = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*;', l):
        inst.append(pend); pend = None
last = None
for i, x in enumerate(inst):
    if x is None: inst[i] = last
    else: last = x
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
by, ie, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
tot = 0
for k, r in enumerate(data):
    ln = inst[k] if k < len(inst) else None
    s = int(r[ci['# Samples']]); tot += s
    by[ln] += s; ie[ln] += int(r[ci['Instructions Executed']])
    for h in stall:
        v = r[ci[h]]
        if v and v != '0': st[ln][h] += int(v)
src = open(cu).read().split('\n')
base = cu.split('/')[-1]
acc = 0
for n in range(l0, l1 + 1):
    k = (base, n)
    if by[k] == 0 and ie[k] == 0: continue
    acc += by[k]
    why = ",".join("%s:%d" % (h.replace('stall_', ''), 100 * v // max(by[k], 1)) for h, v in st[k].most_common(2))
    print(f"{n:5d} {100*by[k]/tot:5.2f}% inst {ie[k]:>10} {src[n-1].strip()[:90]}  [{why}]")
print("range total %.1f%%" % (100 * acc / tot))
