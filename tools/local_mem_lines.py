#!/usr/bin/env python3
"""Developer tool: which CUDA source lines of one kernel touch local memory (stack frame / spills)?  Static: counts LDL / STL
instructions of the kernel's SASS by the source line nvdisasm attributes them to.
usage: local_mem_lines.py libtcpt.so mangled-name-substring   e.g. k_shadeILi5ELb1E"""
import collections, os, re, subprocess, sys, tempfile
lib, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
want, cur, n, chain = False, ("?", 0), collections.Counter(), []
for ln in sass:
    if ln.startswith(".text."):
        want = pat in ln
        continue
    if not want:
        continue
    s = ln.strip()
    m = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
    if m:
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))   # innermost first, then the inlined-at chain
        continue
    if re.match(r"/\*[0-9a-f]+\*/", s):
        if chain:
            cur = chain[0]
        chain = []
    m = re.search(r"\b(LDL|STL)(\.\w+)*\b", s)
    if m:
        n[(cur[0], cur[1], m.group(1))] += 1
root = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc")
for (f, line, op), c in sorted(n.items(), key=lambda kv: -kv[1])[:50]:
    try:
        text = open(os.path.join(root, f)).read().splitlines()[line - 1].strip()[:110]
    except Exception:
        text = ""
    print(f"{c:4d} {op} {f}:{line}  {text}")
