"""Join an ncu --page source --csv dump (SASS rows) with nvdisasm -gi line info: samples per CUDA source line.
usage: ncu_lines.py <ncu_source.csv> <nvdisasm.txt> <kernel-substring> <file.cu> [top]
Also prints the split of samples by stall reason for the top lines."""
import collections, csv, re, sys

src_csv, dis, kname, cu = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l][0]
inst, pend = [], None
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if pend is None:
            pend = (m.group(1).split('/')[-1], int(m.group(2)))  # innermost frame comes first
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*;', l):
        inst.append(pend)
        pend = None
# instructions without their own annotation inherit the previous one
last = None
for i, x in enumerate(inst):
    if x is None:
        inst[i] = last
    else:
        last = x
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_')]
by, ie, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
tot = 0
for k, r in enumerate(data):
    ln = inst[k] if k < len(inst) else None
    s = int(r[ci['# Samples']])
    by[ln] += s; tot += s
    ie[ln] += int(r[ci['Instructions Executed']])
    for h in stall:
        v = r[ci[h]]
        if v and v != '0':
            st[ln][h] += int(v)
src = open(cu).read().split('\n')
print("instructions", len(data), "annotated", len(inst), "samples", tot, "warp-insts", sum(ie.values()))
for ln, s in by.most_common(top):
    txt = src[ln[1] - 1].strip()[:80] if ln and ln[0] == cu.split('/')[-1] else ''
    why = ",".join("%s:%d" % (h.replace('stall_', ''), 100 * v // max(s, 1)) for h, v in st[ln].most_common(3))
    print(f"{100*s/tot:5.1f}% inst {ie[ln]:>11} {str(ln):22} {txt}   [{why}]")
