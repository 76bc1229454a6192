#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel launch in an .ncu-rep (read on the CPU box).

  python scripts/ncu_lines.py <report.ncu-rep> <kernel-regex> <launch-skip> <lib.so> [top]

ncu's CSV source page only exports SASS rows, so the SASS rows are joined, in order, with the line-info annotations
nvdisasm prints for the same function of the in-tree library (needs the .so the profile was taken with)."""
import csv, io, os, re, subprocess, sys, tempfile

rep, kre, skip, so = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
sass, seen = [], set()
for r in rows[2:]:
    if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); sass.append(r)
# demangled -> find the function in the disassembly by matching the base name and the instruction count
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
base = re.search(r"(\w+_kernel)", kname).group(1)
best = None
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, line, funcs = None, 0, {}
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = m.group(1); funcs[cur] = []; continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1), int(m.group(2))); continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            funcs[cur].append(line)
    for fn, lst in funcs.items():
        if base in fn and len(lst) == len(sass):
            best = (fn, lst)
if best is None:
    sys.exit(f"no function containing {base} with {len(sass)} instructions")
agg = {}
tot_i = tot_s = 0
for r, ln in zip(sass, best[1]):
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    a = agg.setdefault(ln, [0, 0, {}]); a[0] += n; a[1] += s; tot_i += n; tot_s += s
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            v = int(r[ix[k]] or 0)
            if v: a[2][k] = a[2].get(k, 0) + v
print(kname[:150]); print(f"instructions {tot_i}  samples {tot_s}")
src_cache = {}
for ln, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    stalls = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{ln[0]}:{ln[1]:<5d} inst {100*n/max(tot_i,1):5.1f}%  samples {100*s/max(tot_s,1):5.1f}%  {stalls}")
