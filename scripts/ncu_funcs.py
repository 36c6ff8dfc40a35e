"""Per-subroutine totals of an `ncu --page source --csv --print-source cuda,sass` export of one kernel:
warp instructions, active lanes per instruction, stall samples.  Subroutine address ranges come from nvdisasm.
usage: python scripts/ncu_funcs.py <source.csv> <lib.so> <kernel-name-substring>"""
import collections, csv, os, re, subprocess, sys, tempfile
path, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", cub], cwd=td, capture_output=True, text=True).stdout
insec = False; labels = []; cursub = "main"; pending = None
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", line)
    if m:
        insec = pat in m.group(1); continue
    if not insec: continue
    if re.match(r"\s*\.section", line): insec = False; continue
    m = re.match(r"\s*\$\S+\$(_Z\w+):", line)
    if m: pending = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+\S", line)
    if m and pending:
        labels.append((int(m.group(1), 16), pending)); pending = None
labels.sort()
rows = list(csv.reader(open(path)))
hdr = None; seen = {}
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 5 or not r[2].startswith("0x"): continue
    a = int(r[2], 16)
    if a in seen: continue
    g = lambda name: int(r[hdr.index(name)] or 0)
    seen[a] = (g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples"), r[3])
base = min(seen)
def sub_of(off):
    name = "main"
    for o, n in labels:
        if off >= o: name = n
    return name
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for a, (i, t, s, sass) in seen.items():
    v = agg[sub_of(a - base)]
    v[0] += i; v[1] += t; v[2] += s; v[3] += 1
ti = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("total warp inst %d, thread inst %d, lanes %.2f, samples %d" % (ti, sum(v[1] for v in agg.values()), sum(v[1] for v in agg.values()) / ti, ts))
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-90s sass %5d inst %5.1f%% lanes %5.1f samples %5.1f%%" % (n[:90], v[3], 100. * v[0] / ti, v[1] / max(v[0], 1), 100. * v[2] / ts))
