"""Where a kernel's warp-instructions and stall samples go, by source file / line, from an .ncu-rep captured with
--import-source on (+ -lineinfo build).  Read here, without a GPU.
usage: python scripts/ncu_regions.py <rep> <cubin-substring> <kernel-substring> [lines-per-file]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, cub, key = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 12
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("PN_LIB", os.path.join(root, "pienerf_b200", "lib", "libpienerf_b200.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cub in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
amap, cur, infn = {}, None, False
for ln in sass.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        infn = key in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if "inlined at" not in ln:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        amap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + key.split("IL")[0]], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex, ith = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base, agg, tot, totex = None, {}, 0, 0
for r in rows[hi + 1:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break                                            # first kernel instance only
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    s, ex, th = int(r[isamp] or 0), int(r[iex] or 0), int(r[ith] or 0)
    loc = amap.get(a - base, ("?", 0))
    g = agg.setdefault(loc, [0, 0, 0])
    g[0] += s; g[1] += ex; g[2] += th
    tot += s; totex += ex
print(f"total: {totex / 1e6:.1f}M warp-instructions, {tot} stall samples")
byfile = {}
for (f, l), (s, ex, th) in agg.items():
    g = byfile.setdefault(f, [0, 0, 0]); g[0] += s; g[1] += ex; g[2] += th
for f, (s, ex, th) in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    if ex < totex * 0.002:
        continue
    print(f"== {f:28s} {ex / 1e6:8.1f}M instr ({100 * ex / totex:4.1f}%)  samples {100 * s / tot:4.1f}%  avg active lanes {th / max(ex, 1):4.1f}")
    lines = sorted(((l, v) for (ff, l), v in agg.items() if ff == f), key=lambda kv: -max(kv[1][1] / totex, kv[1][0] / tot))[:topn]
    for l, (s, ex, th) in sorted(lines):
        print(f"     line {l:4d}: {ex / 1e6:7.1f}M instr {100 * ex / totex:4.1f}%  samples {100 * s / tot:4.1f}%  active {th / max(ex, 1):4.1f}")
