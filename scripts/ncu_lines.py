"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line."""
import csv, collections, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
cur = None; hdr = None
agg = {}; src = {}
def f(x):
    try: return int(x)
    except: return 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iS = hdr.index("# Samples")
        iN = hdr.index("stall_no_inst"); iL = hdr.index("stall_long_sb"); iW = hdr.index("stall_wait"); iB = hdr.index("stall_branch_resolving"); iSS = hdr.index("stall_short_sb")
        continue
    if hdr is None or cur is None or r[0] == "" or r[2] != "-": continue
    key = (cur, int(r[0])); src[key] = r[1]
    agg[key] = [f(r[iI]), f(r[iT]), f(r[iS]), f(r[iN]), f(r[iL]), f(r[iW]), f(r[iB]), f(r[iSS])]
tot = [sum(v[i] for v in agg.values()) for i in range(8)]
print("total inst %d threadinst %d avg lanes %.2f samples %d noinst %d long_sb %d wait %d branch %d short_sb %d" % (tot[0], tot[1], tot[1] / max(tot[0], 1), *tot[2:]))
byf = collections.defaultdict(lambda: [0] * 8)
for (fn, l), v in agg.items():
    for i in range(8): byf[fn][i] += v[i]
for fn, v in byf.items(): print("%-16s inst %11d lanes %5.1f samples %8d noinst %8d" % (fn, v[0], v[1] / max(v[0], 1), v[2], v[3]))
print("--- top lines by samples")
for (k, v) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-14s %4d inst=%9d lanes=%5.1f samp=%6d noinst=%6d lsb=%5d | %s" % (k[0], k[1], v[0], v[1] / max(v[0], 1), v[2], v[3], v[4], src[k].strip()[:100]))
