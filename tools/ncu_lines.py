#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS rows) of a report with the line table of
the matching cubin (`nvdisasm -g`).  Usage: tools/ncu_lines.py REPORT.ncu-rep CUBIN MANGLED_KERNEL_PREFIX UNITS [min_pct]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, kern, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
min_pct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text." + kern)][0]
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith("\t.section"))
cur, table = None, {}
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        table[int(m.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr = rows[1]
c = {k: i for i, k in enumerate(hdr)}
base = int(rows[2][0], 16)
samples, insts, total = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    key = table.get(int(r[0], 16) - base)
    s = int(r[c["# Samples"]] or 0)
    samples[key] += s
    insts[key] += int(r[c["Instructions Executed"]] or 0)
    total += s
cache = {}
print(f"total samples {total}")
for key in sorted(samples, key=lambda k: k or ("", 0)):
    if samples[key] < total * min_pct / 100:
        continue
    f, n = key or ("?", 0)
    if f not in cache:
        try:
            cache[f] = open(f).read().split("\n")
        except OSError:
            cache[f] = []
    text = cache[f][n - 1].strip()[:100] if 0 < n <= len(cache[f]) else ""
    print(f"{f.split('/')[-1]}:{n:4d} {100 * samples[key] / total:5.1f}%  {insts[key] / units:7.1f} inst/unit  {text}")
