"""Warp instructions and stall samples of a kernel in an .ncu-rep, attributed to the FUNCTION of the outermost
source frame (beam.cu member functions) and to the innermost source lines.
usage: ncu_regions.py <rep.ncu-rep> <lib.so> <kernel-mangled-substring> <file.cu> [steps_total]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, so, kname, cu = sys.argv[1:5]
steps = float(sys.argv[5]) if len(sys.argv) > 5 else 444 * 5572.0
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(so)} >/dev/null 2>&1", shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith(os.path.basename(cu).split('.')[0] + '.') and f.endswith('.cubin')][0]
dis = subprocess.run(f"nvdisasm -gi {tmp}/{cub}", shell=True, capture_output=True, text=True).stdout.split('\n')
srccsv = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and kname in l][0]
inst, cur = [], []
for l in dis[start + 1:]:
    if l.startswith('//--------------------- .text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*;', l):
        inst.append(cur if cur else None); cur = []
last = None
for i, x in enumerate(inst):
    if x is None: inst[i] = last
    else: last = x
rows = list(csv.reader(srccsv.split('\n')))
rows = [r for r in rows if r]
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
src = open(cu).read().split('\n')
# function regions of the .cu: "__device__ ... name(" at member indentation
funcs = []
for i, l in enumerate(src):
    m = re.match(r'\s*(?:template.*)?__device__.*?\b(\w+)\s*\(', l)
    if m and not l.strip().startswith('//'): funcs.append((i + 1, m.group(1)))
m2 = [(i + 1, 'run_item') for i, l in enumerate(src) if 'Engine<MODEL>::run_item' in l]
funcs = sorted(funcs + m2)
def fn(n):
    name = '?'
    for ln, nm in funcs:
        if ln <= n: name = nm
        else: break
    return name
base = os.path.basename(cu)
ie, smp, inner = collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
print("sass rows", len(data), "disasm", len(inst))
for k, r in enumerate(data):
    ln = inst[k] if k < len(inst) else None
    n = int(r[ci['Instructions Executed']]); s = int(r[ci['# Samples']])
    tot += n; tots += s
    outer = None
    for fr in (ln or []):
        if fr[0] == base: outer = fr
    key = fn(outer[1]) if outer else 'other'
    ie[key] += n; smp[key] += s
    if ln: inner[ln[0]] += n
print(f"total warp-inst {tot:.4g}  per step {tot/steps:.0f}")
for k, v in ie.most_common(24):
    print(f"{100*v/tot:5.1f}% inst {v/steps:8.0f}/step   samples {100*smp[k]/max(tots,1):5.1f}%  {k}")
print('--- innermost lines')
for k, v in inner.most_common(30):
    print(f"{100*v/tot:5.1f}% {v/steps:7.0f}/step {k} {src[k[1]-1].strip()[:80] if k[0]==base else ''}")
# --- stall samples by the source line of the OUTERMOST frame inside the .cu (where the time is waited for)
osmp, oinst, why = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
stall = [h for h in hdr if h.startswith('stall_')]
for k, r in enumerate(data):
    ln = inst[k] if k < len(inst) else None
    outer = None
    for fr in (ln or []):
        if fr[0] == base: outer = fr
    s = int(r[ci['# Samples']]); osmp[outer] += s; oinst[outer] += int(r[ci['Instructions Executed']])
    for h in stall:
        v = r[ci[h]]
        if v and v != '0': why[outer][h.replace('stall_', '')] += int(v)
print('--- samples by outermost line')
for k, v in osmp.most_common(int(os.environ.get('TOP', '45'))):
    w = ",".join("%s:%d" % (h, 100 * c // max(v, 1)) for h, c in why[k].most_common(3))
    print(f"{100*v/max(tots,1):5.1f}% smp {oinst[k]/steps:6.0f} inst/step {fn(k[1]) if k else '-':18} L{k[1] if k else 0:<5} {src[k[1]-1].strip()[:70] if k else ''}  [{w}]")
