"""Where a kernel touches local memory (LDL/STL): counts per source line (development aid).
usage: python scripts/sass_local.py <lib.so> <kernel-name-substring>"""
import collections, os, re, subprocess, sys, tempfile
lib, pat = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], cwd=td, capture_output=True, text=True).stdout
insec = False; cur = None; sub = "main"; cnt = collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", line)
    if m: insec = pat in m.group(1); sub = "main"; continue
    if not insec: continue
    m = re.match(r"\s*\$\S+\$(_Z\w+):", line)
    if m: sub = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.search(r"\b(LDL|STL)\b", line): cnt[(sub[:30], cur)] += 1
src = {}
for (sb, k), v in sorted(cnt.items(), key=lambda kv: (kv[0][0], kv[0][1] or ("", 0))):
    text = ""
    if k:
        path = os.path.join("abeille_b200/csrc", k[0])
        if os.path.exists(path):
            if path not in src: src[path] = open(path).read().splitlines()
            text = src[path][k[1] - 1].strip()[:90]
    print(v, sb, k, "|", text)
