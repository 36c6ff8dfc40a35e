"""Static + dynamic footprint per source line of an `ncu --page source --csv --print-source cuda,sass` export:
SASS instructions mapped to the line (all / executed), warp instructions, lanes, no-instruction stall samples."""
import csv, collections, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path)))
cur = None; hdr = None; line = None
st = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])  # sass, sass executed, inst, thread inst, samples, no_inst
src = {}
seen = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iS = hdr.index("# Samples"); iN = hdr.index("stall_no_inst"); continue
    if hdr is None: continue
    if r[0] != "":
        line = (cur, int(r[0])); src[line] = r[1]; continue
    if not r[2].startswith("0x") or r[2] in seen: continue
    seen.add(r[2])
    f = lambda i: int(r[i] or 0)
    v = st[line]; v[0] += 1; v[1] += 1 if f(iI) > 0 else 0; v[2] += f(iI); v[3] += f(iT); v[4] += f(iS); v[5] += f(iN)
tot = [sum(v[i] for v in st.values()) for i in range(6)]
print("sass %d (%.1f KB) executed %d (%.1f KB) inst %d lanes %.2f samples %d no_inst %d" % (tot[0], tot[0] / 64., tot[1], tot[1] / 64., tot[2], tot[3] / max(tot[2], 1), tot[4], tot[5]))
byf = collections.defaultdict(lambda: [0] * 6)
for (fn, l), v in st.items():
    for i in range(6): byf[fn][i] += v[i]
for fn, v in sorted(byf.items(), key=lambda kv: -kv[1][2]):
    print("%-28s sass %5d exec %5d inst %5.1f%% lanes %5.1f samples %5.1f%% no_inst %5.1f%%" % (fn, v[0], v[1], 100. * v[2] / tot[2], v[3] / max(v[2], 1), 100. * v[4] / tot[4], 100. * v[5] / max(tot[5], 1)))
key = {"inst": 2, "sass": 1, "samples": 4, "noinst": 5}[sys.argv[3] if len(sys.argv) > 3 else "inst"]
print("--- top lines by", key)
for k, v in sorted(st.items(), key=lambda kv: -kv[1][key])[:top]:
    print("%-14s %4d sass=%4d/%4d inst=%5.2f%% lanes=%5.1f samp=%5.2f%% noinst=%5.2f%% | %s" % (k[0][:14], k[1], v[1], v[0], 100. * v[2] / tot[2], v[3] / max(v[2], 1), 100. * v[4] / tot[4], 100. * v[5] / max(tot[5], 1), src[k].strip()[:90]))
