"""Join an ncu SASS-page CSV (per-instruction stall samples) with nvdisasm line info -> hot source lines.
usage: python scripts/ncu_lines.py <src_page.csv> <nvdisasm -g -c output> <mangled-name-substring> [top]"""
import csv
import re
import sys

src_csv, sass, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file, line) for the chosen function; inline chains: keep the innermost (last) annotation
amap, cur, infn = {}, None, False
for ln in open(sass):
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
        amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
agg, tot, totex = {}, 0, 0
per_inst = []
for r in rows[hi + 1:]:
    if len(r) <= isamp or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    s = int(r[isamp] or 0); ex = int(r[iex] or 0)
    loc, txt = amap.get(a - base, (("?", 0), ""))
    agg.setdefault(loc, [0, 0]); agg[loc][0] += s; agg[loc][1] += ex
    tot += s; totex += ex
    per_inst.append((s, a - base, loc, r[1][:70], ex))
print(f"total samples {tot}, warp-instructions {totex}")
byfile = {}
for (f, l), (s, ex) in agg.items():
    byfile.setdefault(f, [0, 0]); byfile[f][0] += s; byfile[f][1] += ex
for f, (s, ex) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} samples {100 * s / tot:5.1f}%  instr {100 * ex / totex:5.1f}%")
print("-- hottest source lines")
for (f, l), (s, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * s / tot:5.1f}% samp {100 * ex / totex:5.1f}% instr  {f}:{l}")
print("-- hottest instructions")
for s, a, loc, txt, ex in sorted(per_inst, reverse=True)[:15]:
    print(f"{100 * s / tot:5.1f}% {a:06x} {loc[0]}:{loc[1]} {txt}")
