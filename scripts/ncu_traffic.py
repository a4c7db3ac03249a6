"""DRAM traffic of the first-pass field-kernel launch from an `ncu --set full` capture -> profiles/ncu_traffic.json (read by bench.py).
usage: python scripts/ncu_traffic.py <config> <rep.ncu-rep> <rows of that launch> [note]"""
import csv
import json
import os
import subprocess
import sys

config, rep, rows = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
tab = list(csv.reader(out.splitlines()))
hdr, units = tab[0], tab[1]
ik = hdr.index("Kernel Name")
r = next(r for r in tab[2:] if "wave_field_ws_kernel" in r[ik])


def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)


d = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
     "l2_bytes": val("lts__t_bytes.sum") if "lts__t_bytes.sum" in hdr else None, "time_ns": val("gpu__time_duration.sum"), "rows": rows,
     "algorithmic_bytes": rows * 1036, "source": f"profiles/{os.path.basename(rep)} (ncu --set full --clock-control none, first wavefront pass)"}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
allc = json.load(open(p)) if os.path.exists(p) else {}
allc[config] = d
json.dump(allc, open(p, "w"), indent=1)
print(json.dumps(d))
