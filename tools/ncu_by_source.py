#!/usr/bin/env python3
"""Attribute the warp-stall samples and executed instructions of one profiled launch to CUDA source lines and functions.

ncu's CLI source page is per SASS instruction; this joins it (by instruction offset) with `nvdisasm -gi` line info of the same
cubin, then aggregates by innermost source line and by the function that line sits in.
usage: ncu_by_source.py rep.ncu-rep libtcpt.so launch_index [top_n]"""
import collections, csv, re, subprocess, sys, tempfile, os

rep, lib, launch = sys.argv[1], sys.argv[2], int(sys.argv[3])
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
det = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = det[0]
names = collections.OrderedDict()
for r in det[1:]:
    names.setdefault(r[h.index("ID")], r[h.index("Kernel Name")])
kname = list(names.values())[launch]
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
tables, cur = [], None
for r in src:
    if "Source" in r and "# Samples" in r:
        cur = {"h": r, "rows": []}; tables.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["rows"].append(r)
per = len(tables) // len(names)          # the CLI prints each launch's table `per` times
t = tables[launch * per]
hh = t["h"]
ia, isamp, iex, ith = hh.index("Address"), hh.index("# Samples"), hh.index("Instructions Executed"), hh.index("Thread Instructions Executed")
rows = [(int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), int(r[isamp] or 0), int(r[iex] or 0), int(r[ith] or 0), r[hh.index("Source")]) for r in t["rows"]]
base = rows[0][0]
# nvdisasm of the matching function
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
m = re.match(r"(?:void )?(?:tcpt::)?(\w+)(?:<([^>]*)>)?", kname)
fn = m.group(1)
targs = [re.sub(r"\((?:int|bool)\)", "", a).strip() for a in m.group(2).split(",")] if m.group(2) else []
targ_pat = "".join(r"L[ib]" + a + "E" for a in targs)   # <5, 1> -> ILi5ELb1E in the mangled name
sec, lines_info, chain = None, {}, []
want = None
for ln in sass:
    if ln.startswith(".text."):
        sec = ln.strip().rstrip(":")
        ok = fn in sec and (not targs or re.search("I" + targ_pat + "E", sec))
        want = sec if ok else None
        chain = []
        continue
    if want is None:
        continue
    s = ln.strip()
    if s.startswith("//## File"):
        mm = re.match(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', s)
        if mm:
            chain.append((os.path.basename(mm.group(1)), int(mm.group(2))))
        continue
    mm = re.match(r"/\*([0-9a-f]+)\*/", s)
    if mm:
        if chain:
            cur_chain = chain; chain = []
            last = cur_chain
        lines_info[int(mm.group(1), 16)] = last if "last" in dir() else [("?", 0)]
# function table per file: line -> function name (nearest preceding definition)
def func_table(path):
    out = []
    try:
        for i, l in enumerate(open(path), 1):
            mm = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?__(?:device|global)__[^;(]*?\b(\w+)\s*\(", l)
            if mm and not l.strip().startswith("//"):
                out.append((i, mm.group(1)))
    except OSError:
        pass
    return out
root = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc")
ftabs = {f: func_table(os.path.join(root, f)) for f in os.listdir(root) if f.endswith((".cuh", ".cu"))}
def func_of(f, line):
    best = "?"
    for i, n in ftabs.get(f, []):
        if i <= line:
            best = n
        else:
            break
    return best
by_line, by_func, by_outer = collections.Counter(), collections.Counter(), collections.Counter()
ex_line, ex_func = collections.Counter(), collections.Counter()
tot_s = tot_e = 0
for addr, s, ex, th, text in rows:
    ch = lines_info.get(addr - base)
    if not ch:
        ch = [("?", 0)]
    own = [c for c in ch if c[0] in ftabs] or ch      # innermost line in OUR sources
    f, l = own[0]
    by_line[(f, l)] += s; ex_line[(f, l)] += ex
    by_func[func_of(f, l)] += s; ex_func[func_of(f, l)] += ex
    of, ol = own[-1]
    by_outer[func_of(of, ol)] += s
    tot_s += s; tot_e += ex
print(f"# {kname} launch {launch}: {tot_s} samples, {tot_e} warp instructions")
print("## by function (innermost line in our sources): samples%  instr%")
for k, v in by_func.most_common(topn):
    print(f"   {100 * v / max(tot_s, 1):5.1f}%  {100 * ex_func[k] / max(tot_e, 1):5.1f}%  {k}")
print("## by outermost function (non-inlined frame)")
for k, v in by_outer.most_common(15):
    print(f"   {100 * v / max(tot_s, 1):5.1f}%  {k}")
print("## by line")
srcs = {}
for (f, l), v in by_line.most_common(topn):
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(root, f)).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:110] if 0 < l <= len(srcs[f]) else ""
    print(f"   {100 * v / max(tot_s, 1):5.1f}%  {100 * ex_line[(f, l)] / max(tot_e, 1):5.1f}%  {f}:{l}  {text}")
