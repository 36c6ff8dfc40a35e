"""SASS instruction counts per source line of one kernel (development aid).
usage: python scripts/sass_lines.py <lib.so> <kernel-name-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], cwd=td, capture_output=True, text=True).stdout
cnt = collections.Counter(); sub = collections.Counter(); only = sys.argv[4] if len(sys.argv) > 4 else None
insec = False; cur = None; cursub = "main"
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", line)
    if m:
        insec = pat in m.group(1); cur = None; cursub = "main"; continue
    if line.startswith("\t.section") or line.startswith(".section"):
        insec = False
    if not insec: continue
    m = re.match(r"\s*\$\S+\$(_Z\w+):", line)
    if m: cursub = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", line):
        sub[cursub] += 1
        if only is None or only in cursub: cnt[cur] += 1
byfile = collections.Counter()
for k, c in cnt.items(): byfile[k[0] if k else None] += c
print("total", sum(cnt.values())); print("by subroutine", sub.most_common()); print("by file", byfile.most_common())
for k, c in cnt.most_common(top): print(k, c)
