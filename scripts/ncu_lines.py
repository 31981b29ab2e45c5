"""Per-source-line hot spots of one kernel from an ncu report (--import-source on, -lineinfo).
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [top]"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: collections.Counter())
src = {}
hdr = None
cur_file = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        line = int(d["Line No"])
    except ValueError:
        continue
    key = (cur_file, line)
    cols = list(zip(hdr, r))
    src[key] = r[1]
    a = agg[key]
    a["samples"] += int(d.get("# Samples") or 0)
    a["inst"] += int(d.get("Instructions Executed") or 0)
    for name, v in cols:
        if name.startswith("stall_") and "Not Issued" not in name and v:
            a[name] += int(v)
tot = sum(a["samples"] for a in agg.values()) or 1
toti = sum(a["inst"] for a in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print(f"{100 * a['samples'] / tot:5.1f}% smp {100 * a['inst'] / toti:5.1f}% inst  {key[0]}:{key[1]:<4d} {' '.join(f'{k}={v}' for k, v in st):40s} | {src[key].strip()[:110]}")
