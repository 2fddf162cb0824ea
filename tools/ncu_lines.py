"""Join an ncu SASS-level source page with nvdisasm line info: per source line instruction / stall-sample shares.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    in_k, line = False, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
        if m:
            in_k = kern in m.group(1)
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and line:
            addr2line[int(m.group(1), 16)] = line
csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csv_txt.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg, base = {}, None
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    base = a if base is None else base
    key = addr2line.get(a - base, ("?", 0))
    d = agg.setdefault(key, [0, 0, {}])
    d[0] += int(r[ie] or 0)
    d[1] += int(r[isamp] or 0)
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            d[2][hdr[c]] = d[2].get(hdr[c], 0) + v
ti, ts = sum(d[0] for d in agg.values()), sum(d[1] for d in agg.values())
print("total warp instructions %d, samples %d" % (ti, ts))
src = {}
for (f, l), d in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in src:
        p = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.dirname(os.path.abspath(lib)) + "/..") if f in fs]
        src[f] = open(p[0]).read().splitlines() if p else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    stalls = ",".join("%s:%d" % (k.replace("stall_", ""), v) for k, v in sorted(d[2].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% inst %5.1f%% samp  %s:%d  [%s]  %s" % (100.0 * d[0] / max(ti, 1), 100.0 * d[1] / max(ts, 1), f, l, stalls, text))

# ---- summary: stall reasons over the whole kernel and instruction share per source region -------------------
tot = {}
for d in agg.values():
    for k, v in d[2].items():
        tot[k] = tot.get(k, 0) + v
print("stall reasons:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(ts, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
if len(sys.argv) > 5:
    # regions given as name:lo-hi,...  (line ranges of the main source file)
    for spec in sys.argv[5].split(","):
        name, rng = spec.split(":")
        lo, hi_ = [int(x) for x in rng.split("-")]
        i = sum(d[0] for (f, l), d in agg.items() if f == "physics.cu" and lo <= l <= hi_)
        s_ = sum(d[1] for (f, l), d in agg.items() if f == "physics.cu" and lo <= l <= hi_)
        print("region %-12s lines %d-%d: %.1f%% inst, %.1f%% samples" % (name, lo, hi_, 100.0 * i / max(ti, 1), 100.0 * s_ / max(ts, 1)))
    other = sum(d[1] for (f, l), d in agg.items() if f != "physics.cu")
    print("other files (intrinsics headers etc.): %.1f%% samples" % (100.0 * other / max(ts, 1)))
